"""The weight-stationary, halo-tile trunk kernel (csrc/conv_ws.cu; fp16 operands, one MMA per product) against
a float64 convolution of the very fp16 values it consumed, and against the tap-streaming CTA-pair kernel it
replaces for the fp16-operand passes (src/autoencoder_imgcomp.py:229-234,257-262,275-288)."""
import numpy as np
import pytest
import torch

from oracle import dsin_oracle as O

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _case(n, hh, ww, nres, act, seed=0):
    from dsin_b200 import ops
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 128, hh, ww)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 128, 128)) / np.sqrt(9 * 128)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 128).astype(np.float32)
    shift = rng.standard_normal(128).astype(np.float32)
    layer = ops.ConvLayer(w, scale, shift, act=act)
    tcl = ops.ConvTC(layer)
    xh = _nhwc(torch.tensor(x).cuda()).half()
    res = [_nhwc(torch.tensor(rng.standard_normal((n, 128, hh, ww)).astype(np.float32)).cuda()).half()
           for _ in range(nres)]
    return ops, tcl, w, scale, shift, xh, res


def _ref64(tcl, w, scale, shift, xh, res, act):
    """float64 result from the fp16 operands the kernel reads: x as fp16, weights as their packed fp16 hi plane."""
    wscale = (torch.tensor(scale).cuda() / tcl.scale).cpu().double()          # the per-cout power of two
    w_hi = tcl.w_hi.cpu().double().reshape(3, 3, 128, 128).permute(0, 1, 3, 2)  # [ky][kx][cin][cout], scaled
    xq = xh.permute(0, 3, 1, 2).cpu().double()
    y = O.conv2d_same(xq, w_hi.numpy()) / wscale.view(1, -1, 1, 1)
    y = y * torch.tensor(scale).double().view(1, -1, 1, 1) + torch.tensor(shift).double().view(1, -1, 1, 1)
    if act:
        y = torch.relu(y)
    for r in res:
        y = y + r.permute(0, 3, 1, 2).cpu().double()
    return y


@pytest.mark.parametrize("shape", [(2, 20, 36), (1, 80, 306), (3, 9, 17), (1, 16, 24), (2, 33, 70)])
@pytest.mark.parametrize("nres", [0, 1, 2])
def test_conv_ws_matches_float64_and_streaming_kernel(shape, nres):
    n, hh, ww = shape
    act = 1 if nres < 2 else 0
    ops, tcl, w, scale, shift, xh, res = _case(n, hh, ww, nres, act, seed=hh + nres)
    r1 = (res[0], None) if nres > 0 else None
    r2 = (res[1], None) if nres > 1 else None
    y_ws, lo = ops.conv_tc((xh, None), tcl, res1=r1, res2=r2, terms=1)
    assert lo is None and y_ws.dtype == torch.float16
    y_st, _ = ops.conv_tc((xh, None), tcl, res1=r1, res2=r2, terms=1, flags=ops.CONV_NO_WEIGHT_STATIONARY)
    ref = _ref64(tcl, w, scale, shift, xh, res, act)
    got = y_ws.permute(0, 3, 1, 2).cpu().double()
    old = y_st.permute(0, 3, 1, 2).cpu().double()
    # fp32 accumulation of exact products, then ONE rounding to fp16 (relative 2^-11): the two kernels differ only in
    # summation order (fp32), so they agree to one fp16 ulp; against float64 the bound is half an ulp + fp32 noise
    tol = 2.0 ** -10 * torch.clamp(ref.abs(), min=1.0)
    assert bool(((got - ref).abs() <= tol).all()), float(((got - ref).abs() / tol).max())
    assert bool(((got - old).abs() <= 2 * tol).all())
    frac_equal = float((got == old).double().mean())
    assert frac_equal > 0.99, frac_equal


def test_conv_ws_is_the_default_for_fp16_operand_trunk_layers():
    """The dispatch: terms = 1, 3x3, 128 -> 128 runs conv_ws_kernel (one more launch on the handle, same result whether
    the residual is given or not) and a lo residual plane falls back to the streaming kernel."""
    ops, tcl, w, scale, shift, xh, res = _case(1, 32, 48, 1, 1, seed=5)
    y0, _ = ops.conv_tc((xh, None), tcl, res1=(res[0], None), terms=1)
    lo = torch.zeros_like(res[0])
    y1, _ = ops.conv_tc((xh, None), tcl, res1=(res[0], lo), terms=1)  # lo plane present -> streaming kernel
    assert float((y0.float() - y1.float()).abs().max()) <= 2.0 ** -9 * float(y0.float().abs().max())


# ----------------------------------------------------------------------------- fp32-class halo kernel (conv_h3.cu)
@pytest.mark.parametrize("shape", [(2, 20, 36), (1, 80, 306), (3, 9, 17), (2, 33, 70)])
@pytest.mark.parametrize("nres", [0, 1, 2])
def test_conv_h3_matches_float64_and_is_closer_than_the_single_accumulator_kernel(shape, nres):
    """3-term layer: halo-tile kernel with separate accumulators for the large and the small product terms vs a float64
    convolution of the split operands it consumed, and vs the tap-streaming kernel (all terms in one accumulator)."""
    from dsin_b200 import ops
    n, hh, ww = shape
    rng = np.random.default_rng(100 + hh + nres)
    x = rng.standard_normal((n, 128, hh, ww)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 128, 128)) / np.sqrt(9 * 128)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 128).astype(np.float32)
    shift = rng.standard_normal(128).astype(np.float32)
    act = ops.ACT_RELU if nres < 2 else ops.ACT_NONE
    tcl = ops.ConvTC(ops.ConvLayer(w, scale, shift, act=act))
    xs = ops.f32_to_split(_nhwc(torch.tensor(x).cuda()))
    res = [ops.f32_to_split(_nhwc(torch.tensor(rng.standard_normal((n, 128, hh, ww)).astype(np.float32)).cuda()))
           for _ in range(nres)]
    r1 = res[0] if nres > 0 else None
    r2 = res[1] if nres > 1 else None
    y_h3 = ops.split_to_f32(*ops.conv_tc(xs, tcl, res1=r1, res2=r2, terms=3)).permute(0, 3, 1, 2).cpu().double()
    y_st = ops.split_to_f32(*ops.conv_tc(xs, tcl, res1=r1, res2=r2, terms=3, flags=ops.CONV_NO_HALO))
    y_st = y_st.permute(0, 3, 1, 2).cpu().double()
    xq = ops.split_to_f32(*xs).permute(0, 3, 1, 2).cpu().double()
    ref = O.conv2d_same(xq, w.astype(np.float64))
    ref = ref * torch.tensor(scale).double().view(1, -1, 1, 1) + torch.tensor(shift).double().view(1, -1, 1, 1)
    if act == ops.ACT_RELU:
        ref = torch.relu(ref)
    for r in res:
        ref = ref + ops.split_to_f32(*r).permute(0, 3, 1, 2).cpu().double()
    e_h3, e_st = (y_h3 - ref).abs(), (y_st - ref).abs()
    scale_ref = max(1.0, float(ref.abs().max()))
    print("3-term %s nres %d: halo kernel max %.2e rms %.2e | single accumulator max %.2e rms %.2e"
          % (shape, nres, float(e_h3.max()), float(e_h3.pow(2).mean().sqrt()), float(e_st.max()),
             float(e_st.pow(2).mean().sqrt())))
    assert float(e_h3.max()) < 1e-5 * scale_ref
    assert float(e_st.max()) < 1e-5 * scale_ref
    assert float(e_h3.pow(2).mean().sqrt()) <= 1.05 * float(e_st.pow(2).mean().sqrt())


# ----------------------------------------------------------------------------- 32-channel halo kernel (conv_h32.cu)
@pytest.mark.parametrize("terms,tol", [(3, 1e-5), (1, 4e-3)])
@pytest.mark.parametrize("dil", [1, 2, 4])
@pytest.mark.parametrize("shape", [(2, 40, 48), (1, 37, 52), (1, 320, 1224)])
def test_conv_h32_sinet_layers_match_float64_and_streaming_kernel(shape, dil, terms, tol):
    """SI-Net 3x3 32->32 layers with dilation 1/2/4, bias + LeakyReLU (src/siNet.py:9-10,31-34,39): halo-tile kernel vs
    float64 on the operands it consumed and vs the tap-streaming kernel."""
    from dsin_b200 import ops
    n, hh, ww = shape
    rng = np.random.default_rng(7 * dil + hh)
    x = rng.standard_normal((n, 32, hh, ww)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 32, 32)) / np.sqrt(9 * 32)).astype(np.float32)
    bias = (0.3 * rng.standard_normal(32)).astype(np.float32)
    tcl = ops.ConvTC(ops.ConvLayer(w, None, bias, dilation=dil, act=ops.ACT_LRELU02))
    xs = ops.f32_to_split(_nhwc(torch.tensor(x).cuda()), with_lo=terms == 3)
    got = ops.conv_tc(xs, tcl, terms=terms)
    old = ops.conv_tc(xs, tcl, terms=terms, flags=ops.CONV_NO_HALO)
    got = ops.split_to_f32(*got).permute(0, 3, 1, 2).cpu().double()
    old = ops.split_to_f32(*old).permute(0, 3, 1, 2).cpu().double()
    xq = ops.split_to_f32(*xs).permute(0, 3, 1, 2).cpu().double()
    ref = O.conv2d_same(xq, w.astype(np.float64), dilation=dil) + torch.tensor(bias).double().view(1, -1, 1, 1)
    ref = torch.maximum(0.2 * ref, ref)
    scale_ref = max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) < tol * scale_ref
    assert float((got - old).abs().max()) < 2 * tol * scale_ref


# ----------------------------------------------------------------------------- 32-channel row-band kernel (conv_dil.cu)
@pytest.mark.parametrize("terms,tol", [(3, 1e-5), (1, 4e-3)])
@pytest.mark.parametrize("shape,dil", [((2, 40, 48), 8), ((1, 37, 56), 16), ((1, 64, 136), 8), ((1, 80, 144), 32),
                                       ((1, 80, 144), 128), ((1, 320, 1224), 8), ((1, 320, 1224), 16),
                                       ((1, 320, 1224), 64), ((2, 320, 1224), 128), ((1, 33, 264), 24)])
def test_conv_dil_sinet_layers_match_float64_and_streaming_kernel(shape, dil, terms, tol):
    """SI-Net 3x3 32->32 layers with dilation >= 8, bias + LeakyReLU (src/siNet.py:9-10,34-38): row-band kernel vs float64
    on the operands it consumed and vs the tap-streaming kernel.  Covers images narrower than one 128-pixel tile, heights
    below the dilation (every vertical neighbour outside the image), chains of one row, segments, and a dilation that is
    not a multiple of 8."""
    from dsin_b200 import ops
    n, hh, ww = shape
    rng = np.random.default_rng(11 * dil + hh)
    x = rng.standard_normal((n, 32, hh, ww)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 32, 32)) / np.sqrt(9 * 32)).astype(np.float32)
    bias = (0.3 * rng.standard_normal(32)).astype(np.float32)
    tcl = ops.ConvTC(ops.ConvLayer(w, None, bias, dilation=dil, act=ops.ACT_LRELU02))
    xs = ops.f32_to_split(_nhwc(torch.tensor(x).cuda()), with_lo=terms == 3)
    l0 = ops.launch_count()
    got = ops.conv_tc(xs, tcl, terms=terms)
    assert ops.launch_count() - l0 == 1
    old = ops.conv_tc(xs, tcl, terms=terms, flags=ops.CONV_NO_HALO)
    got = ops.split_to_f32(*got).permute(0, 3, 1, 2).cpu().double()
    old = ops.split_to_f32(*old).permute(0, 3, 1, 2).cpu().double()
    xq = ops.split_to_f32(*xs).permute(0, 3, 1, 2).cpu().double()
    ref = O.conv2d_same(xq, w.astype(np.float64), dilation=dil) + torch.tensor(bias).double().view(1, -1, 1, 1)
    ref = torch.maximum(0.2 * ref, ref)
    scale_ref = max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) < tol * scale_ref
    assert float((got - old).abs().max()) < 2 * tol * scale_ref
