"""GPU: the fused encoder block (block_fused.cu: conv1 -> control FiLM -> SiLU -> conv2 -> ResidualConnection as ONE
kernel, hidden tensor kept in shared memory; reference model.py:773-778 + 507-511 + 789-792) against

  * the same block run as the two separate kernels (conv1 mode 0, conv2 mode 1) on the same 16-bit operands -- the two
    paths round the hidden tensor identically and differ only in fp32 accumulation order;
  * a CPU fp32 reference (torch.nn.functional) computed on the same rounded operands;

over ragged widths around the 126-pixel strip / 252-pixel strip pair, single rows, segment heights that exercise the
ring wrap and the junk rows above and below a segment, several rounds per CTA pair (max_ctas), both operand types, with
and without FiLM; and the whole model with fused blocks against the oracle and against the unfused schedule."""
import pytest
import torch
from torch.nn import functional as F

from oracle import make_oracle, max_abs_err, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _operands(shape, seed, dt, with_film=True):
    g = torch.Generator().manual_seed(seed)
    B, H, W = shape
    zf0 = torch.randn(B, H, W, 48, generator=g)
    zb = zf0.to(dt)
    w1 = torch.randn(96, 48, 3, 3, generator=g) / (3.0 * 48 ** 0.5)
    w2 = torch.randn(48, 96, 3, 3, generator=g) / (3.0 * 96 ** 0.5)
    film = None
    if with_film:
        film = torch.empty(B, 2, 96)
        film[:, 0] = 1 + 0.3 * torch.randn(B, 96, generator=g)
        film[:, 1] = 0.3 * torch.randn(B, 96, generator=g)      # a non-zero shift: SiLU(shift) != 0 outside the image
    return zf0, zb, w1, w2, film


def _reference(zf0, zb, w1, w2, film, dt):
    """fp32 math on the 16-bit operands, the hidden tensor rounded to 16 bits as the kernels do."""
    acc1 = F.conv2d(zb.float().permute(0, 3, 1, 2), w1.to(dt).float(), padding=1)
    if film is not None:
        acc1 = acc1 * film[:, 0][:, :, None, None] + film[:, 1][:, :, None, None]
    hid = F.silu(acc1).to(dt).float()
    acc2 = F.conv2d(hid, w2.to(dt).float(), padding=1).permute(0, 2, 3, 1)
    return zf0 + acc2


CASES = [
    # (B, H, W), seg_rows, max_ctas
    ((1, 1, 1), 0, 0), ((1, 1, 126), 0, 0), ((1, 2, 127), 0, 0), ((1, 3, 130), 0, 0), ((1, 5, 252), 0, 0),
    ((1, 4, 253), 0, 0), ((2, 9, 300), 0, 0), ((1, 23, 260), 8, 0), ((1, 23, 260), 3, 2), ((1, 40, 140), 1, 2),
    ((2, 17, 400), 5, 4), ((1, 64, 520), 0, 6), ((3, 7, 100), 2, 2), ((1, 33, 960), 0, 0),
]


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,seg_rows,max_ctas", CASES)
def test_fused_block_matches_two_kernels_and_reference(dev, shape, seg_rows, max_ctas, dt):
    from ultrazoom_b200 import ops

    zf0, zb, w1, w2, film = _operands(shape, sum(shape) + seg_rows, dt)
    w1p, w2p = ops.pack_conv_weight(w1, dev, dtype=dt), ops.pack_conv_weight(w2, dev, dtype=dt)
    fd = film.to(dev)
    # two kernels
    hid = ops.conv3x3(zb.to(dev), w1p, 0, fd, use_tc=True)
    zf_two = zf0.to(dev).contiguous()
    zb_two = ops.conv3x3(hid, w2p, 1, None, zf_two, use_tc=True)
    # one kernel
    zf_one = zf0.to(dev).contiguous()
    zb_one = ops.block_fused(zb.to(dev), w1p, w2p, fd, zf_one, seg_rows=seg_rows, max_ctas=max_ctas)
    torch.cuda.synchronize()
    ref = _reference(zf0, zb, w1, w2, film, dt)
    ulp = 2.0 ** -11 if dt == torch.float16 else 2.0 ** -8
    # hidden values ~O(1) rounded to 16 bits: an accumulation-order flip of one hidden element moves an output by
    # ~ulp * |w2|; the fp32 sums themselves differ by ~1e-6
    tol = 3e-3 if dt == torch.float16 else 2e-2
    assert (zf_one.cpu() - ref).abs().max().item() <= tol, (zf_one.cpu() - ref).abs().max().item()
    assert (zf_one - zf_two).abs().max().item() <= tol
    assert torch.equal(zb_one, zf_one.to(dt))                           # the shadow is exactly round16(zf)
    assert (zb_one.float() - zb_two.float()).abs().max().item() <= tol + 4 * ulp * ref.abs().max().item()
    # deterministic, and bit-identical however the image is cut into segments and rounds (the accumulator block of a
    # row -- hence its fp32 association -- is tied to the image row, not to the work decomposition)
    zf_again = zf0.to(dev).contiguous()
    ops.block_fused(zb.to(dev), w1p, w2p, fd, zf_again, seg_rows=seg_rows, max_ctas=max_ctas)
    assert torch.equal(zf_again, zf_one)
    zf_cut = zf0.to(dev).contiguous()
    ops.block_fused(zb.to(dev), w1p, w2p, fd, zf_cut, seg_rows=max(1, shape[1] // 2), max_ctas=2)
    assert torch.equal(zf_cut, zf_one)


def test_fused_block_without_film_and_input_untouched(dev):
    from ultrazoom_b200 import ops

    dt = torch.float16
    zf0, zb, w1, w2, _ = _operands((2, 11, 270), 5, dt, with_film=False)
    w1p, w2p = ops.pack_conv_weight(w1, dev, dtype=dt), ops.pack_conv_weight(w2, dev, dtype=dt)
    zbd = zb.to(dev)
    keep = zbd.clone()
    zf = zf0.to(dev).contiguous()
    out = ops.block_fused(zbd, w1p, w2p, None, zf)
    assert torch.equal(zbd, keep)                                       # the input stream is only read
    ref = _reference(zf0, zb, w1, w2, None, dt)
    assert (zf.cpu() - ref).abs().max().item() <= 3e-3 and torch.equal(out, zf.to(dt))
    with pytest.raises(AssertionError):
        ops.block_fused(zb.to(dev)[..., :32].contiguous(), w1p, w2p, None, zf)


@pytest.mark.parametrize("name,shape", [("MewZoom-2X-Ctrl", (2, 3, 70, 300)), ("MewZoom-2X", (1, 3, 129, 127))])
def test_model_with_fused_blocks(dev, name, shape):
    """The 2X models run their encoder as fused blocks by default: against the oracle (stated 2X tolerance), against
    the unfused schedule (tune block=2) and the SIMT twin, float and 8-bit images, CUDA-graph replay."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom, _native

    o = make_oracle(name, seed=0)
    m = MewZoom(**MODEL_CONFIGS[name])
    m.load_state_dict(o.state_dict())
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(31)
    x = torch.rand(shape, generator=g)
    c = torch.rand(shape[0], 3, generator=g) if o.control_features else None
    cd = None if c is None else c.to(dev)
    eng = m._engine(dev)
    assert eng.lib.mz_model_fused_block(eng.handle) == 1
    y = m.upscale(x.to(dev), cd)
    ref = o.upscale(x, c)
    assert max_abs_err(y.cpu(), ref) <= 4e-3 and psnr(y.cpu(), ref) >= 64.0, (max_abs_err(y.cpu(), ref), psnr(y.cpu(), ref))
    assert torch.equal(y, m.upscale(x.to(dev), cd))
    graph = m.capture(x.to(dev), cd)
    assert torch.equal(graph.replay(), y)
    y8 = m.upscale((x * 255).to(torch.uint8).to(dev), cd)
    assert y8.dtype == torch.uint8
    m.set_conv_tune(0, dev, block=2)                                    # two kernels per block
    assert eng.lib.mz_model_fused_block(eng.handle) == 0
    y2 = m.upscale(x.to(dev), cd)
    assert (y - y2).abs().max().item() <= 2e-3
    m._flags_extra = _native.FLAG_SIMT_CONV
    ys = m.upscale(x.to(dev), cd)
    m._flags_extra = 0
    assert (y - ys).abs().max().item() <= 2e-3
    m.set_conv_tune(0, dev, block=0, seg_rows=7, max_ctas=4)            # many segments, several rounds per CTA pair
    y3 = m.upscale(x.to(dev), cd)
    assert torch.equal(y, y3)                                           # bit-identical for any work decomposition
    assert torch.equal(m.upscale(x.to(dev)[:1], None if cd is None else cd[:1]), y[:1])   # batch independence
    m3 = MewZoom(**MODEL_CONFIGS["MewZoom-3X"]).to(dev)                 # other channel counts keep two kernels per block
    e3 = m3._engine(dev)
    assert e3.lib.mz_model_fused_block(e3.handle) == 0
    m3.set_conv_tune(0, dev, block=1)
    with pytest.raises(RuntimeError, match="fused encoder block was required"):
        m3.upscale(torch.rand(1, 3, 8, 8, device=dev))
