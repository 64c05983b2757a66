"""GPU parity proper: MewZoom (C ABI, sm_100a kernels) vs the oracle and vs fixtures made by the reference's own
leaf classes.  BASELINE.json envelope: max-abs <= 2e-2 on [0,1] pixels, PSNR >= 45 dB.  Stated per config, for the
default fp16 tensor-core operands (fp32 accumulation, fp32 residual stream), un-clamped forward AND clamped upscale,
per-image random control vectors:
    2X (48 ch / 20 layers):  max-abs <= 4e-3, PSNR >= 64 dB
    3X (54 ch / 30 layers):  max-abs <= 6e-3, PSNR >= 62 dB
    4X (96 ch / 40 layers):  max-abs <= 8e-3, PSNR >= 60 dB
and for operand_dtype="bfloat16" (the type BASELINE.json's north_star names; 8x coarser mantissa), README control
vector (0.5, 0.2, 0.3):   max-abs <= 2e-2 (2X, 3X) / 2.5e-2 (4X), PSNR >= 50 dB on these test sizes.  The 40-layer
bf16 variant sits ON the 2e-2 envelope (CPU emulation: 0.0206 at 256x256 for 4X-Ctrl; 0.0196..0.0204 here depending
on the fp32 accumulation order of the chosen kernel configuration), which is why fp16 is the default."""
import pytest
import torch

from oracle import make_oracle, max_abs_err, psnr, residual_rms
from tests.helpers import CASES, load_case, oracle_from_case

pytestmark = pytest.mark.gpu

TOL = {"MewZoom-2X": (4e-3, 64.0), "MewZoom-3X": (6e-3, 62.0), "MewZoom-4X": (8e-3, 60.0)}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def _model_from(cfg, sd, dev, operand_dtype="float16"):
    from ultrazoom_b200 import MewZoom

    m = MewZoom(**cfg, operand_dtype=operand_dtype)
    m.load_state_dict(sd)
    return m.to(dev).eval()


@pytest.mark.parametrize("name", CASES)
def test_reference_fixtures(dev, name):
    cfg, sd, x, c, out = load_case(name)
    m = _model_from(cfg, sd, dev)
    cd = c.to(dev) if c is not None else None
    y = m.forward(x.to(dev), cd).cpu()
    yc = m.upscale(x.to(dev), cd).cpu()
    assert max_abs_err(y, out["forward"]) <= 2e-3 and max_abs_err(yc, out["upscale"]) <= 2e-3
    assert psnr(yc, out["upscale"]) >= 70.0
    assert float(yc.min()) >= 0.0 and float(yc.max()) <= 1.0          # reference tests/test_model.py:161-169
    # SIMT twin and the diagnostic halo mode give the same answer within accumulate-order noise
    from ultrazoom_b200 import _native

    m._flags_extra = _native.FLAG_SIMT_CONV
    ys = m.forward(x.to(dev), cd).cpu()
    m._flags_extra = _native.FLAG_SKIP_FROM_BUFFER
    yb = m.forward(x.to(dev), cd).cpu()
    m._flags_extra = 0
    assert max_abs_err(ys, y) <= 1e-3 and max_abs_err(yb, y) <= 1e-5
    m.set_conv_tune(-1, dev, halo_mode=1)
    assert max_abs_err(m.forward(x.to(dev), cd).cpu(), y) <= 1e-3
    # bf16 operands: the coarser variant still reproduces the reference fixture within the envelope
    mb = _model_from(cfg, sd, dev, "bfloat16")
    assert max_abs_err(mb.upscale(x.to(dev), cd).cpu(), out["upscale"]) <= 1e-2


@pytest.mark.parametrize("name,shape", [
    ("MewZoom-2X", (1, 3, 64, 200)), ("MewZoom-2X-Ctrl", (2, 3, 96, 128)),
    ("MewZoom-3X-Ctrl", (2, 3, 45, 131)), ("MewZoom-4X", (1, 3, 33, 97)), ("MewZoom-4X-Ctrl", (1, 3, 64, 160)),
])
def test_named_models_against_oracle(dev, name, shape):
    o = make_oracle(name, seed=0)
    m = _model_from(dict(upscale_ratio=o.upscale_ratio, num_channels=o.num_channels, hidden_ratio=o.hidden_ratio,
                         num_encoder_layers=o.num_encoder_layers, control_features=o.control_features),
                    o.state_dict(), dev)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(shape, generator=g)
    c = torch.rand(shape[0], 3, generator=g) if o.control_features else None
    assert residual_rms(o, x, c) >= 0.1                                 # degeneracy guard
    tol_abs, tol_psnr = TOL[name.replace("-Ctrl", "")]
    cd = c.to(dev) if c is not None else None
    with torch.inference_mode():
        ref_f, ref_u = o.forward(x, c), o.upscale(x, c)
    got_f, got_u = m.forward(x.to(dev), cd).cpu(), m.upscale(x.to(dev), cd).cpu()
    assert max_abs_err(got_f, ref_f) <= tol_abs, max_abs_err(got_f, ref_f)
    assert max_abs_err(got_u, ref_u) <= tol_abs
    assert psnr(got_u, ref_u) >= tol_psnr, psnr(got_u, ref_u)
    if o.control_features:
        mb = _model_from(dict(upscale_ratio=o.upscale_ratio, num_channels=o.num_channels, hidden_ratio=o.hidden_ratio,
                              num_encoder_layers=o.num_encoder_layers, control_features=o.control_features),
                         o.state_dict(), dev, "bfloat16")
        cr = torch.tensor([[0.5, 0.2, 0.3]])
        ref_b = o.upscale(x, cr)
        got_b = mb.upscale(x.to(dev), cr.to(dev)).cpu()
        tol_b = 2.5e-2 if name.startswith("MewZoom-4X") else 2e-2
        assert max_abs_err(got_b, ref_b) <= tol_b and psnr(got_b, ref_b) >= 50.0


@pytest.mark.parametrize("name", ["MewZoom-2X-Ctrl", "MewZoom-3X-Ctrl", "MewZoom-4X-Ctrl"])
def test_ragged_shapes_against_the_simt_twin(dev, name):
    """Edge and ragged shapes (single row / column, widths around the 128-pixel tile, odd batches): the tcgen05 path,
    whatever configuration the launcher picks (resident bank, CTA pairs, row count, dependent launch), agrees with
    the SIMT direct convolution running the same 16-bit operands and epilogues."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom, _native

    torch.manual_seed(5)
    cfg = dict(MODEL_CONFIGS[name])
    cfg["num_encoder_layers"] = 3            # the kernels and their chaining are what is under test here
    m = MewZoom(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(6)
    shapes = [(1, 1, 1), (1, 1, 130), (1, 2, 127), (3, 3, 129), (1, 5, 1), (2, 4, 256), (1, 7, 257), (1, 31, 64),
              (5, 2, 33), (1, 65, 200)]
    for B, H, W in shapes:
        x, c = torch.rand(B, 3, H, W, generator=g).to(dev), torch.rand(B, 3, generator=g).to(dev)
        m._flags_extra = 0
        y = m.upscale(x, c)
        m._flags_extra = _native.FLAG_SIMT_CONV
        ys = m.upscale(x, c)
        m._flags_extra = 0
        assert tuple(y.shape) == (B, 3, H * cfg["upscale_ratio"], W * cfg["upscale_ratio"])
        assert (y - ys).abs().max().item() <= 2e-3, (name, (B, H, W), (y - ys).abs().max().item())
        assert torch.equal(y, m.upscale(x, c))


def test_uint8_image_io(dev):
    """8-bit images in and out (the callers' decode_image / ToDtype(scale=True) before and save_image / ToPILImage after
    the reference's upscale, test_compare.py:53-57,89, README.md:81) against the oracle run on x8 / 255."""
    o = make_oracle("MewZoom-2X-Ctrl", seed=2)
    m = _model_from(dict(upscale_ratio=2, num_channels=48, hidden_ratio=2, num_encoder_layers=20, control_features=3),
                    o.state_dict(), dev)
    g = torch.Generator().manual_seed(8)
    x8 = torch.randint(0, 256, (2, 3, 37, 150), generator=g, dtype=torch.uint8)
    c = torch.rand(2, 3, generator=g)
    with torch.inference_mode():
        ref = o.upscale(x8.float() / 255.0, c)
    y8 = m.upscale(x8.to(dev), c.to(dev)).cpu()
    assert y8.dtype == torch.uint8 and tuple(y8.shape) == (2, 3, 74, 300)
    want = torch.floor(ref * 255.0 + 0.5).clamp(0, 255).to(torch.uint8)          # save_image's rounding
    diff = (y8.int() - want.int()).abs()
    assert int(diff.max()) <= 1 and float((diff == 0).float().mean()) >= 0.9     # 4e-3 envelope = one 8-bit step
    # against our own fp32 path the 8-bit result is the same rounding of (numerically almost) the same value
    yf = m.upscale((x8.float() / 255.0).to(dev), c.to(dev)).cpu()
    d2 = (y8.int() - torch.floor(yf * 255.0 + 0.5).int()).abs()
    assert int(d2.max()) <= 1 and float((d2 == 0).float().mean()) >= 0.999
    m.u8_truncate = True                                                          # ToPILImage's pic.mul(255).byte()
    yt = m.upscale(x8.to(dev), c.to(dev)).cpu()
    d3 = (yt.int() - torch.floor(yf * 255.0).clamp(0, 255).int()).abs()
    assert int(d3.max()) <= 1 and float((d3 == 0).float().mean()) >= 0.999
    m.u8_truncate = False
    # host buffers (pinned, two lanes) and every ratio
    h8 = m.upscale_host(x8, c)
    assert torch.equal(h8, y8)
    with pytest.raises(AssertionError):
        m.forward(x8.to(dev), c.to(dev))
    for name, r in (("MewZoom-3X-Ctrl", 3), ("MewZoom-4X-Ctrl", 4)):
        o2 = make_oracle(name, seed=3)
        m2 = _model_from(dict(upscale_ratio=o2.upscale_ratio, num_channels=o2.num_channels, hidden_ratio=o2.hidden_ratio,
                              num_encoder_layers=o2.num_encoder_layers, control_features=3), o2.state_dict(), dev)
        xs = torch.randint(0, 256, (1, 3, 20, 131), generator=g, dtype=torch.uint8)
        cs = torch.rand(1, 3, generator=g)
        with torch.inference_mode():
            w2 = torch.floor(o2.upscale(xs.float() / 255.0, cs) * 255.0 + 0.5).clamp(0, 255).to(torch.uint8)
        got = m2.upscale(xs.to(dev), cs.to(dev)).cpu()
        d = (got.int() - w2.int()).abs()
        assert int(d.max()) <= 2 and float((d <= 1).float().mean()) >= 0.999, (name, int(d.max()))


def test_control_vector_broadcast_and_api(dev):
    from ultrazoom_b200 import ControlVector, ONNXModel

    o = make_oracle("MewZoom-2X-Ctrl", seed=1)
    m = _model_from(dict(upscale_ratio=2, num_channels=48, hidden_ratio=2, num_encoder_layers=20, control_features=3),
                    o.state_dict(), dev)
    x = torch.rand(2, 3, 24, 40, generator=torch.Generator().manual_seed(3))
    c = ControlVector(gaussian_blur=0.5, gaussian_noise=0.2, jpeg_compression=0.3).to_tensor()   # README.md:118-122
    ref = o.upscale(x, c)
    a = m.upscale(x.to(dev), c.to(dev)).cpu()                            # (3,) broadcast
    b = m.upscale(x.to(dev), c.repeat(2, 1).to(dev)).cpu()               # (B,3) as validate.py:73-94
    assert torch.equal(a, b)
    assert max_abs_err(a, ref) <= 4e-3
    assert max_abs_err(ONNXModel(m)(x.to(dev), c.to(dev)).cpu(), a) == 0.0
    h = m.upscale_host(x, c)                                             # host-buffer entry point
    assert max_abs_err(h, a) == 0.0
    # a batch of 5 is pipelined in five one-image chunks over the two lanes, per-image control vectors sliced per chunk
    g5 = torch.Generator().manual_seed(5)
    x5, c5 = torch.rand(5, 3, 24, 40, generator=g5), torch.rand(5, 3, generator=g5)
    assert max_abs_err(m.upscale_host(x5, c5), m.upscale(x5.to(dev), c5.to(dev)).cpu()) == 0.0
    # frame-stream form: frames alternate between the two lanes, each result valid after its lane's wait
    xs = [torch.rand(1, 3, 24, 40, generator=g5).pin_memory() for _ in range(4)]
    outs = [torch.empty(1, 3, 48, 80).pin_memory() for _ in range(4)]
    for i, xi in enumerate(xs):
        if i >= 2:
            m.host_wait(i & 1)
        m.upscale_host(xi, c, out=outs[i], lane=i & 1)
    m.host_wait(-1)
    for xi, oi in zip(xs, outs):
        assert max_abs_err(oi, m.upscale(xi.to(dev), c.to(dev)).cpu()) == 0.0
    with pytest.raises(AssertionError):
        m.upscale(x.to(dev), torch.rand(3, 3, device=dev))
    with pytest.raises(AssertionError):
        m.upscale(x.to(dev), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.upscale(x, c)


def test_size_independent_properties_at_full_size(dev):
    """BASELINE configs[1] frame size (960x540 -> 1920x1080), checked through properties the domain offers:
    batch independence (each image's result does not depend on its batch mates), determinism, range, and
    tiling exactness against the un-tiled result (halo 2L+1)."""
    from ultrazoom_b200.sharding import upscale_tiled

    o = make_oracle("MewZoom-2X-Ctrl", seed=0)
    m = _model_from(dict(upscale_ratio=2, num_channels=48, hidden_ratio=2, num_encoder_layers=20, control_features=3),
                    o.state_dict(), dev)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 540, 960, generator=g).to(dev)
    c = torch.rand(2, 3, generator=g).to(dev)
    y = m.upscale(x, c)
    assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    assert torch.equal(y, m.upscale(x, c))                               # deterministic
    y1 = m.upscale(x[1:], c[1:])
    assert torch.equal(y[1:], y1)                                        # batch independence, bit-exact
    tiled = upscale_tiled(m.upscale, x[:1], c[:1], 2, 20, rows=2, cols=2)
    assert (tiled - y[:1]).abs().max().item() <= 1e-5                    # same kernels, different patch origins
    # a crop of the big frame agrees with the CPU oracle on the same crop + halo
    R = 41
    crop = x[:1, :, :64 + R, :64 + R].cpu()
    ref = o.upscale(crop, c[:1].cpu())[:, :, :128, :128]
    assert max_abs_err(y[:1, :, :128, :128].cpu(), ref) <= 4e-3


@pytest.mark.parametrize("name,shape", [("MewZoom-2X-Ctrl", (2, 3, 40, 56)), ("MewZoom-4X-Ctrl", (1, 3, 33, 130)),
                                        ("MewZoom-3X", (1, 3, 24, 40))])
def test_cuda_graph_replay_matches_the_eager_call(dev, name, shape):
    """MewZoom.capture: the whole call recorded into a CUDA graph (cluster launches, dependent-launch edges) and replayed
    on NEW inputs copied into its buffers is bit-identical to the ordinary call, float and uint8 images, and still
    within the oracle's tolerance."""
    from ultrazoom_b200 import MODEL_CONFIGS

    o = make_oracle(name, seed=11)
    m = _model_from(MODEL_CONFIGS[name], o.state_dict(), dev)
    g = torch.Generator().manual_seed(12)
    ctrl = MODEL_CONFIGS[name]["control_features"] > 0
    x0, x1 = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    c0 = torch.rand(shape[0], 3, generator=g) if ctrl else None
    c1 = torch.rand(shape[0], 3, generator=g) if ctrl else None
    to = lambda t: None if t is None else t.to(dev)  # noqa: E731
    graph = m.capture(to(x0), to(c0))
    assert torch.equal(graph.replay(), m.upscale(to(x0), to(c0)))
    y1 = graph(to(x1), to(c1)).clone()
    assert torch.equal(y1, m.upscale(to(x1), to(c1)))
    tol, min_psnr = TOL[name.replace("-Ctrl", "")]
    ref = o.upscale(x1, c1)
    assert max_abs_err(y1.cpu(), ref) <= tol and psnr(y1.cpu(), ref) >= min_psnr
    for _ in range(3):                                                   # replays are repeatable
        assert torch.equal(graph.replay(), y1)
    fwd = m.capture(to(x0), to(c0), clamp=False)                         # un-clamped forward
    assert torch.equal(fwd(to(x1), to(c1)), m.forward(to(x1), to(c1)))
    x8 = (x1 * 255).to(torch.uint8).to(dev)                              # 8-bit images in and out
    g8 = m.capture(x8, to(c1))
    assert g8.y.dtype == torch.uint8 and torch.equal(g8.replay(), m.upscale(x8, to(c1)))
    with pytest.raises(AssertionError):
        graph(to(x1)[:, :, :-1], to(c1))


@pytest.mark.parametrize("cfg,tol", [
    (dict(upscale_ratio=2, num_channels=96, hidden_ratio=4, num_encoder_layers=3, control_features=3), 4e-3),   # 384 = 2 x 192
    (dict(upscale_ratio=3, num_channels=80, hidden_ratio=4, num_encoder_layers=2, control_features=0), 4e-3),   # 320 = 2 x 160
    (dict(upscale_ratio=2, num_channels=128, hidden_ratio=4, num_encoder_layers=2, control_features=3), 6e-3),  # 512 = 2 x 256
    (dict(upscale_ratio=4, num_channels=48, hidden_ratio=1, num_encoder_layers=3, control_features=3), 4e-3),   # hidden_ratio 1
    # above 128 channels conv2 is sliced as well (each launch adds into its channel slice of the residual stream)
    (dict(upscale_ratio=2, num_channels=160, hidden_ratio=2, num_encoder_layers=2, control_features=3), 6e-3),  # 2 x 80 / 2 x 160
    (dict(upscale_ratio=2, num_channels=256, hidden_ratio=1, num_encoder_layers=2, control_features=0), 6e-3),  # 2 x 128 / 256
    (dict(upscale_ratio=3, num_channels=384, hidden_ratio=2, num_encoder_layers=2, control_features=3), 8e-3),  # 3 x 128 / 3 x 256
])
def test_hidden_widths_beyond_one_umma_tile(dev, cfg, tol):
    """InvertedBottleneck allows hidden_ratio 1, 2 and 4 and any channel count (reference model.py:737-738; the 0.3.0
    U-Net runs it at 192 and 384 channels).  A hidden width above the 256 columns of one UMMA tile runs conv1 as slices
    of output channels (each with its own filter bank and FiLM rows), more than 128 channels slice conv2 the same way;
    the result matches the oracle like any other model, the SIMT twin, and is deterministic."""
    from ultrazoom_b200 import _native

    o = make_oracle(cfg, seed=21)
    m = _model_from(cfg, o.state_dict(), dev)
    g = torch.Generator().manual_seed(22)
    x = torch.rand(2, 3, 37, 150, generator=g)
    c = torch.rand(2, 3, generator=g) if cfg["control_features"] else None
    ref = o.upscale(x, c)
    assert residual_rms(o, x, c) >= 0.05
    cd = None if c is None else c.to(dev)
    y = m.upscale(x.to(dev), cd)
    assert max_abs_err(y.cpu(), ref) <= tol, max_abs_err(y.cpu(), ref)
    assert psnr(y.cpu(), ref) >= 60.0
    assert torch.equal(y, m.upscale(x.to(dev), cd))
    m._flags_extra = _native.FLAG_SIMT_CONV
    ys = m.upscale(x.to(dev), cd)
    m._flags_extra = 0
    assert (y - ys).abs().max().item() <= 2e-3


@pytest.mark.parametrize("cfg", [
    dict(upscale_ratio=3, num_channels=54, hidden_ratio=2, num_encoder_layers=3, control_features=3),    # 56 / 112 of 64 / 128
    dict(upscale_ratio=2, num_channels=80, hidden_ratio=2, num_encoder_layers=2, control_features=0),    # 80 / 160 of 96 / 160
    dict(upscale_ratio=4, num_channels=20, hidden_ratio=4, num_encoder_layers=2, control_features=3),    # 24 / 80 of 32 / 96
])
def test_dense_layouts_are_bit_identical(dev, cfg, monkeypatch):
    """A tensor may hold fewer channels in memory than its GEMM is wide (ConvArgs::in_extent, EpiParams::out_extent /
    zf_extent: the tensor maps end at the extent, the zero padding exists in shared memory only).  Every combination --
    fp32 stream, 16-bit shadow, hidden tensor (MZ_DENSE_LAYOUT bits 1 | 2 | 4; the default is the fp32 stream only) --
    gives the padded layout's result bit for bit, on the tcgen05 path and on the SIMT twin, and the workspace shrinks."""
    import ctypes as C

    from ultrazoom_b200 import _native

    o = make_oracle(cfg, seed=31)
    g = torch.Generator().manual_seed(32)
    x = torch.rand(2, 3, 41, 150, generator=g).to(dev)
    c = torch.rand(2, 3, generator=g).to(dev) if cfg["control_features"] else None
    outs, sizes = {}, {}
    for mask in (0, 1, 2, 4, 7):
        monkeypatch.setenv("MZ_DENSE_LAYOUT", str(mask))
        m = _model_from(cfg, o.state_dict(), dev)
        outs[mask] = m.upscale(x, c)
        eng = m._engine(dev)
        need = C.c_size_t()
        _native.check(eng.lib.mz_workspace_bytes(eng.handle, 2, 41, 150, C.byref(need)))
        sizes[mask] = need.value
        m._flags_extra = _native.FLAG_SIMT_CONV
        ys = m.upscale(x, c)
        m._flags_extra = 0
        assert (outs[mask] - ys).abs().max().item() <= 2e-3, mask
        if mask:
            assert torch.equal(outs[mask], outs[0]), (mask, (outs[mask] - outs[0]).abs().max().item())
    assert max_abs_err(outs[0].cpu(), o.upscale(x.cpu(), None if c is None else c.cpu())) <= 6e-3
    assert sizes[7] < sizes[0] and sizes[7] <= sizes[1] < sizes[0] and sizes[2] < sizes[0] and sizes[4] <= sizes[0]


def test_skipped_zero_k_steps_change_nothing(dev, monkeypatch):
    """conv2 of the 54-channel model reads a hidden tensor of 108 channels padded to 128: the eighth 16-channel k-step of
    every tap multiplies zeros and is not issued (TcParams::kt_last).  The result equals the full K loop bit for bit."""
    cfg = dict(upscale_ratio=3, num_channels=54, hidden_ratio=2, num_encoder_layers=4, control_features=3)
    o = make_oracle(cfg, seed=41)
    g = torch.Generator().manual_seed(42)
    x = torch.rand(1, 3, 70, 300, generator=g).to(dev)
    c = torch.rand(1, 3, generator=g).to(dev)
    y = _model_from(cfg, o.state_dict(), dev).upscale(x, c)
    monkeypatch.setenv("MZ_NO_KSKIP", "1")
    y_full = _model_from(cfg, o.state_dict(), dev).upscale(x, c)
    assert torch.equal(y, y_full), (y - y_full).abs().max().item()
    assert max_abs_err(y.cpu(), o.upscale(x.cpu(), c.cpu())) <= 6e-3


def test_put_core_assembles_the_frame(dev):
    """Spatial sharding, stitch step: each tile's HR core is put into the assembled frame with 2-D
    copies (mz_put_plane_async) -- into a device buffer (a peer GPU's in the multi-GPU run, tools/tiled_8k.py) or into
    pinned host memory; the assembled frame equals the un-tiled result bit for bit."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom
    from ultrazoom_b200.sharding import halo_radius, plan_tiles, put_core, run_tile

    torch.manual_seed(31)
    cfg = dict(MODEL_CONFIGS["MewZoom-2X-Ctrl"])
    cfg["num_encoder_layers"] = 4
    m = MewZoom(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(32)
    x, c = torch.rand(2, 3, 70, 300, generator=g).to(dev), torch.rand(2, 3, generator=g).to(dev)
    full = m.upscale(x, c)
    plan = plan_tiles(70, 300, 2, 2, halo_radius(4), align_w=128)
    on_gpu = torch.zeros_like(full)
    on_host = torch.zeros(full.shape, dtype=full.dtype).pin_memory()
    for t in plan:
        core = run_tile(m.upscale, x, c, t, 2)
        put_core(on_gpu, core, t, 2)
        put_core(on_host, core, t, 2)
    torch.cuda.synchronize()
    assert torch.equal(on_gpu, full) and torch.equal(on_host, full.cpu())
    from ultrazoom_b200.sharding import run_tile_into, share_frame

    shared = share_frame(tuple(full.shape), torch.float32, 0, 0, dev)    # the IPC-exportable frame (owner side here)
    assert torch.count_nonzero(shared.tensor).item() == 0
    for t in plan[:2]:
        put_core(shared.tensor, run_tile(m.upscale, x, c, t, 2), t, 2)
    for t in plan[2:]:
        run_tile_into(m, x, c, t, 2, shared.tensor)
    torch.cuda.synchronize()
    assert torch.equal(shared.tensor, full)
    shared.close()
    x8 = (x * 255).to(torch.uint8)                                       # 8-bit frames stitch the same way
    full8, frame8 = m.upscale(x8, c), torch.zeros(2, 3, 140, 600, dtype=torch.uint8, device=dev)
    for t in plan:
        put_core(frame8, run_tile(m.upscale, x8, c, t, 2), t, 2)
    torch.cuda.synchronize()
    assert torch.equal(frame8, full8)


@pytest.mark.parametrize("name", ["MewZoom-2X-Ctrl", "MewZoom-3X-Ctrl", "MewZoom-4X-Ctrl"])
def test_windowed_output_assembles_the_frame(dev, name):
    """mz_upscale_window: the head kernel of each haloed tile stores only the tile's core, straight into the assembled
    frame (the multi-GPU run points it at a peer GPU's buffer: tools/tiled_8k.py).  The frame equals the un-tiled
    result bit for bit, float and 8-bit images, tcgen05 and SIMT kernels; nothing outside a window is touched."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom, _native
    from ultrazoom_b200.sharding import halo_radius, plan_tiles, run_tile_into

    torch.manual_seed(41)
    cfg = dict(MODEL_CONFIGS[name])
    cfg["num_encoder_layers"] = 3
    r = cfg["upscale_ratio"]
    m = MewZoom(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(42)
    x, c = torch.rand(2, 3, 45, 290, generator=g).to(dev), torch.rand(2, 3, generator=g).to(dev)
    plan = plan_tiles(45, 290, 2, 2, halo_radius(3), align_w=128)
    for simt in (0, _native.FLAG_SIMT_CONV):
        m._flags_extra = simt
        full = m.upscale(x, c)
        frame = torch.full_like(full, -1.0)
        for t in plan[:-1]:
            run_tile_into(m, x, c, t, r, frame)
        last = plan[-1]
        assert torch.all(frame[:, :, last.y0 * r:last.y1 * r, last.x0 * r:last.x1 * r] == -1.0)   # untouched so far
        run_tile_into(m, x, c, last, r, frame)
        assert torch.equal(frame, full), (name, simt, (frame - full).abs().max().item())
    m._flags_extra = 0
    x8 = (x * 255).to(torch.uint8)
    full8, frame8 = m.upscale(x8, c), torch.zeros(2, 3, 45 * r, 290 * r, dtype=torch.uint8, device=dev)
    for t in plan:
        run_tile_into(m, x8, c, t, r, frame8)
    assert torch.equal(frame8, full8)
    with pytest.raises(AssertionError):                                  # a window that does not fit the frame
        m.upscale_into(x, c, frame, (0, 45, 0, 290), (r, 0))


@pytest.mark.parametrize("name,L,k", [("MewZoom-2X-Ctrl", 5, 1), ("MewZoom-2X-Ctrl", 6, 2), ("MewZoom-4X-Ctrl", 4, 2), ("MewZoom-3X-Ctrl", 3, 3)])
def test_halo_refresh_tiling_is_bit_exact(dev, name, L, k):
    """sharding.upscale_tiled_refresh: tiles carry a halo of only 2k+1 pixels, run k encoder blocks at a time
    (mz_upscale_stage) and refresh their halo ring from the neighbours' cores in between.  All tiles in this process
    (device-to-device copies; the 2-process form is tests/test_gpu_multiprocess.py): the frame equals the un-tiled
    result bit for bit -- fused blocks (2X: the 16-bit stream alternates between two buffers) and two-kernel blocks."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom
    from ultrazoom_b200.sharding import upscale_tiled_refresh

    torch.manual_seed(61)
    cfg = dict(MODEL_CONFIGS[name])
    cfg["num_encoder_layers"] = L
    r = cfg["upscale_ratio"]
    m = MewZoom(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(62)
    x, c = torch.rand(1, 3, 50, 300, generator=g).to(dev), torch.rand(1, 3, generator=g).to(dev)
    full = m.upscale(x, c)
    for rows, cols in ((2, 2), (1, 3)):
        frame = torch.full_like(full, -1.0)
        state = upscale_tiled_refresh(m, x, c, r, L, rows, cols, k, frame, align_w=128)
        assert torch.equal(frame, full), (rows, cols, (frame - full).abs().max().item())
        frame.fill_(-1.0)
        upscale_tiled_refresh(m, x, c, r, L, rows, cols, k, frame, align_w=128, state=state)   # cached plan + workspaces
        assert torch.equal(frame, full)
    x8 = (x * 255).to(torch.uint8)
    full8, frame8 = m.upscale(x8, c), torch.zeros(1, 3, 50 * r, 300 * r, dtype=torch.uint8, device=dev)
    upscale_tiled_refresh(m, x8, c, r, L, 2, 2, k, frame8, align_w=128)
    assert torch.equal(frame8, full8)
