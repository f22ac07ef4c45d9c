"""main.py-style inference entry (/root/reference/src/main.py:101-126,182-224) on libdsin_b200.

    python -m dsin_b200.main [-ae_config PATH] [-pc_config PATH] [--weights CKPT] [--synthetic N | --pairs X Y]

Keeps the reference's argument names and its test loop: build AE with the five injected callables, load
the model (`<cwd>/weights/<load_model_name>/model`, a TF-V2 checkpoint, when the config says `load_model`;
`--weights` overrides it and also accepts an .npz), and for every test pair call ``siNet_get_reconstructed``,
clip, save the PNG and optionally append to the loss lists.  Pairs come from the reference's pair lists
(`<cwd>/data_paths/<file_path_test>` through DataProvider.Dataset), or from ``--synthetic N`` (seeded
generator) or ``--pairs x.npy y.npy`` arrays shaped (N,3,H,W).  ``--validate`` adds the validation step of the
reference's training loop (mean ``siNet_validate`` loss); the training flags in the config are ignored.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import config_parser, synth
from .AE import AE
from .DataProvider import Dataset
from .decoder_imgcomp import decoder
from .encoder_imgcomp import encoder
from .siFinder import siFinder
from .siFull_img import SI_full_img
from .siNet import siNet
from .utils import loss_list_saver, save_test_imgs_fn  # noqa: F401  (re-exported like the reference's `from utils import *`)

_HERE = os.path.dirname(os.path.abspath(__file__))


def get_run_params(args, current_directory=None):
    """src/main.py:182-208; the hard-wired presentation flags become command-line switches."""
    current_directory = current_directory or os.getcwd()
    ae_config, ae_rel = config_parser.parse(args.ae_config_path)
    pc_config, pc_rel = config_parser.parse(args.pc_config_path)
    return {"ae_config": ae_config, "ae_config_rel_path": ae_rel, "pc_config": pc_config,
            "pc_config_rel_path": pc_rel, "batch_size": ae_config.batch_size,
            "root_weights": current_directory + "/weights/",
            "root_save_img": args.save_dir if args.save_dir is not None else current_directory + "/images/",
            "load_model_name": ae_config.load_model_name, "load_model": ae_config.load_model,
            "test_model": ae_config.test_model, "save_test_img": not args.no_save_test_img,
            "create_loss_list": args.create_loss_list}


def main(run_dict, args):
    ae_config = run_dict["ae_config"]
    cur_dir = os.path.join(os.getcwd(), "data_paths") + os.sep
    ae = AE(ae_config, run_dict["pc_config"], encoder, decoder, siFinder, SI_full_img, siNet, cur_dir)
    model_name = "NA"
    if args.weights:
        ae.load_model(args.weights)
        model_name = os.path.splitext(os.path.basename(args.weights))[0]
    elif run_dict["load_model"] and not args.random_init:  # src/main.py:39-41
        model_name = run_dict["load_model_name"]
        ae.load_model(run_dict["root_weights"] + model_name + "/model")
    H, W = ae_config.crop_size
    if args.pairs:
        xs, ys = np.load(args.pairs[0]), np.load(args.pairs[1])
        batches = ([xs[i:i + 1], ys[i:i + 1]] for i in range(xs.shape[0]))
        n_test = xs.shape[0]
    elif args.synthetic:
        xs, ys = synth.make_batch(args.synthetic, H, W, seed=1000)
        batches = ([xs[i:i + 1], ys[i:i + 1]] for i in range(xs.shape[0]))
        n_test = xs.shape[0]
    else:  # src/main.py:34-37,101-104
        data = Dataset(ae_config, cur_dir)
        _val_names, test_names = data.get_data_size()
        n_test = len(test_names)
        batches = (data.get_data_for_test() for _ in range(n_test))
    if args.validate:  # the validation step of the reference's training loop (src/main.py:64-73), on its own
        if args.pairs or args.synthetic:
            val_batches = [[xs[i:i + 1], ys[i:i + 1]] for i in range(xs.shape[0])]
        else:
            val_names, _test_names = data.get_data_size()
            val_batches = (data.get_data_for_val() for _ in range(len(val_names) // run_dict["batch_size"]))
        val_sum, val_iterations = 0.0, 0
        for x_val, y_val in val_batches:
            val_sum += ae.siNet_validate(x_val, y_val)
            val_iterations += 1
        val_loss = val_sum / float(max(val_iterations, 1))
        print("validation loss = {:.6f} over {:d} batches".format(val_loss, val_iterations))
        run_dict["val_loss"] = val_loss
    results = []
    root_save_img = run_dict["root_save_img"]
    if not root_save_img.endswith(os.sep):
        root_save_img += os.sep
    if run_dict["test_model"]:
        for i in range(n_test):
            print("Processing test image number {:d}".format(i))
            x_test, y_test = next(batches)
            y_dec, y_syn, x_dec, x_with_si, bpp = ae.siNet_get_reconstructed(x_test, y_test)
            x_dec = np.clip(x_dec, 0, 255)
            x_with_si = np.clip(x_with_si, 0, 255)
            img_index = 0
            if run_dict["save_test_img"]:
                save_test_imgs_fn(root_save_img, model_name, x_with_si[img_index], i, bpp)
            if run_dict["create_loss_list"]:  # src/main.py:120-126
                os.makedirs(root_save_img, exist_ok=True)
                x_rec = x_with_si
                if np.average(x_rec[img_index]) == 0:  # AE_only: x_with_si is zero -> use x_dec
                    x_rec = x_dec
                loss_list_saver(x_test, y_test, x_rec, y_syn, x_test.shape[0], str(model_name), bpp, root_save_img)
            results.append(float(bpp))
            print("  bpp = {:.5f}".format(bpp))
            if args.real_bpp:  # entropy-code the symbols for real (probclass_imgcomp.py:361: "--real_bpp")
                streams = ae.compress(x_test)
                real = 8.0 * sum(len(b) for b in streams) / (x_test.shape[0] * x_test.shape[2] * x_test.shape[3])
                print("  real bpp = {:.5f} ({} bytes incl. container)".format(real, sum(len(b) for b in streams)))
                if args.save_dir is not None or run_dict["save_test_img"]:
                    path = os.path.join(root_save_img, model_name)
                    os.makedirs(path, exist_ok=True)
                    for k, blob in enumerate(streams):
                        with open(os.path.join(path, "{}_{}.dspc".format(i, k)), "wb") as f:
                            f.write(blob)
    return results


def build_parser():
    parser = argparse.ArgumentParser()
    cfg = os.path.join(_HERE, "run_configs")
    parser.add_argument("-ae_config", "-ae_configs", "--ae_config_path", "--ae_configs_path", type=str,
                        help="AE config file path", default=os.path.join(cfg, "ae_run_configs"))
    parser.add_argument("-pc_config", "-pc_configs", "--pc_config_path", "--pc_configs_path", type=str,
                        help="PC config file path", default=os.path.join(cfg, "pc_run_configs"))
    parser.add_argument("--weights", type=str, default=None,
                        help="TF-V2 checkpoint prefix (.../model) or .npz keyed by TF variable names; "
                             "default: <cwd>/weights/<load_model_name>/model when the config sets load_model")
    parser.add_argument("--random_init", action="store_true", help="skip load_model: seeded random-init weights")
    parser.add_argument("--synthetic", type=int, default=0, help="use N seeded synthetic pairs instead of the pair lists")
    parser.add_argument("--pairs", nargs=2, default=None, help="x.npy y.npy, each (N,3,H,W)")
    parser.add_argument("--save_dir", type=str, default=None, help="image / list output root (default <cwd>/images/)")
    parser.add_argument("--no_save_test_img", action="store_true", help="reference default is to save (main.py:203)")
    parser.add_argument("--create_loss_list", action="store_true", help="append per-image metric lists (main.py:206)")
    parser.add_argument("--validate", action="store_true",
                        help="first report the mean validation loss (AE.siNet_validate, src/main.py:64-73) over the "
                             "validation split (or over the --synthetic / --pairs images)")
    parser.add_argument("--real_bpp", action="store_true",
                        help="also range-code the symbols (PC1 bitstream) and report / save the real size")
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    main(get_run_params(a), a)
