"""main.py-style inference entry (/root/reference/src/main.py:101-126,182-224) on libdsin_b200.

    python -m dsin_b200.main [-ae_config PATH] [-pc_config PATH] [--weights W.npz] [--synthetic N]

Keeps the reference's argument names and its test loop: build AE with the five injected callables, load
the model, and for every test pair call ``siNet_get_reconstructed``, clip, and report bpp.  The KITTI
tf.data pipeline (DataProvider.py) is out of scope; pairs come from ``--synthetic N`` (seeded generator)
or from ``--pairs x.npy y.npy`` arrays shaped (N,3,H,W).  Training flags in the config are ignored.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import config_parser, synth
from .AE import AE
from .decoder_imgcomp import decoder
from .encoder_imgcomp import encoder
from .siFinder import siFinder
from .siFull_img import SI_full_img
from .siNet import siNet

_HERE = os.path.dirname(os.path.abspath(__file__))


def get_run_params(args):
    ae_config, ae_rel = config_parser.parse(args.ae_config_path)
    pc_config, pc_rel = config_parser.parse(args.pc_config_path)
    return {"ae_config": ae_config, "ae_config_rel_path": ae_rel, "pc_config": pc_config,
            "pc_config_rel_path": pc_rel, "batch_size": ae_config.batch_size,
            "load_model_name": ae_config.load_model_name, "load_model": ae_config.load_model,
            "test_model": ae_config.test_model, "save_test_img": args.save_dir is not None}


def save_test_imgs_fn(root_save_img, model_name, x_with_si, i, bpp):
    """PNG writer with the reference's naming and uint8 truncation (src/utils.py:102-111)."""
    from PIL import Image
    path = os.path.join(root_save_img, model_name)
    os.makedirs(path, exist_ok=True)
    img = Image.fromarray(np.transpose(x_with_si, (1, 2, 0)).astype("uint8"), "RGB")
    img.save(os.path.join(path, str(i) + "_" + "{:.5f}bpp.png".format(bpp)))


def main(run_dict, args):
    ae_config = run_dict["ae_config"]
    cur_dir = os.path.join(os.getcwd(), "data_paths") + os.sep
    ae = AE(ae_config, run_dict["pc_config"], encoder, decoder, siFinder, SI_full_img, siNet, cur_dir)
    model_name = "NA"
    if args.weights:
        ae.load_model(args.weights)
        model_name = os.path.splitext(os.path.basename(args.weights))[0]
    H, W = ae_config.crop_size
    if args.pairs:
        xs, ys = np.load(args.pairs[0]), np.load(args.pairs[1])
    else:
        xs, ys = synth.make_batch(args.synthetic, H, W, seed=1000)
    results = []
    if run_dict["test_model"]:
        for i in range(xs.shape[0]):
            print("Processing test image number {:d}".format(i))
            x_test, y_test = xs[i:i + 1], ys[i:i + 1]
            y_dec, y_syn, x_dec, x_with_si, bpp = ae.siNet_get_reconstructed(x_test, y_test)
            x_dec = np.clip(x_dec, 0, 255)
            x_with_si = np.clip(x_with_si, 0, 255)
            if run_dict["save_test_img"]:
                save_test_imgs_fn(args.save_dir, model_name, x_with_si[0], i, bpp)
            results.append(float(bpp))
            print("  bpp = {:.5f}".format(bpp))
    return results


def build_parser():
    parser = argparse.ArgumentParser()
    cfg = os.path.join(_HERE, "run_configs")
    parser.add_argument("-ae_config", "-ae_configs", "--ae_config_path", "--ae_configs_path", type=str,
                        help="AE config file path", default=os.path.join(cfg, "ae_run_configs"))
    parser.add_argument("-pc_config", "-pc_configs", "--pc_config_path", "--pc_configs_path", type=str,
                        help="PC config file path", default=os.path.join(cfg, "pc_run_configs"))
    parser.add_argument("--weights", type=str, default=None, help=".npz keyed by TF variable names")
    parser.add_argument("--synthetic", type=int, default=2, help="number of synthetic pairs")
    parser.add_argument("--pairs", nargs=2, default=None, help="x.npy y.npy, each (N,3,H,W)")
    parser.add_argument("--save_dir", type=str, default=None, help="write PNGs like the reference")
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    main(get_run_params(a), a)
