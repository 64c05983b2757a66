"""Small-shape run of every kernel variant behind the C ABI, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(tools/sanitize.sh; SURVEY.md section 5 "race detection"): bicubic r = 2/3/4, stem, FiLM table, device weight packer, the
tcgen05 convolution in every epilogue MODE (0 conv1+FiLM+SiLU, 1 conv2+residual, 2 head) and VAR (0 plain,
1 CTA pair / cta_group::2, 2 vertical taps fused along N), resident and streamed filter banks, 4 and 8 epilogue warps,
the SIMT twin, and whole small models through mz_upscale (dependent launches included).  Every result is also checked
against the SIMT twin so that a sanitizer-induced timing change that exposes a protocol bug shows up as a mismatch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import MODEL_CONFIGS, MewZoom, _native, ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    n = 0
    for r in (2, 3, 4):
        ops.bicubic(torch.rand(1, 3, 9, 37, generator=g).to(dev), r)
        n += 1
    cases = [  # cin, cout, (B,H,W), tune
        (48, 96, (1, 5, 150), {}), (48, 96, (1, 5, 150), dict(fuse=2)), (48, 96, (1, 5, 150), dict(resident=2)),
        (96, 48, (1, 6, 140), {}), (96, 48, (1, 6, 140), dict(fuse=2)), (96, 192, (1, 4, 260), {}),
        (96, 192, (1, 4, 260), dict(pair=1)), (96, 192, (1, 4, 260), dict(pair=2, resident=2)),
        (192, 96, (1, 4, 260), {}), (192, 96, (1, 4, 260), dict(pair=2)), (64, 128, (1, 3, 130), dict(epi_warps=4)),
        (128, 64, (2, 3, 130), dict(epi_warps=4)), (32, 48, (1, 7, 130), dict(fuse=1, rows=4)),
    ]
    for cin, cout, (B, H, W), tune in cases:
        cin_p, cout_p = ops.padded_channels(cin), ops.padded_channels(cout)
        inp = torch.zeros(B, H, W, cin_p, dtype=torch.float16)
        inp[..., :cin] = torch.randn(B, H, W, cin, generator=g).half()
        w = torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)
        wp = ops.pack_conv_weight(w, dev)
        film = torch.ones(B, 2, cout_p)
        film[:, 1] = 0.1
        t = _native.tune(**tune)
        a = ops.conv3x3(inp.to(dev), wp, 0, film.to(dev), use_tc=True, tune=t)
        b = ops.conv3x3(inp.to(dev), wp, 0, film.to(dev), use_tc=False)
        assert (a.float() - b.float()).abs().max().item() <= 2e-2, ("mode 0", cin, cout, tune)
        if cout_p <= 128:
            zf0 = torch.randn(B, H, W, cout_p, generator=g)
            zfa, zfb = zf0.to(dev).contiguous(), zf0.to(dev).contiguous()
            ops.conv3x3(inp.to(dev), wp, 1, None, zfa, use_tc=True, tune=t)
            ops.conv3x3(inp.to(dev), wp, 1, None, zfb, use_tc=False)
            assert (zfa - zfb).abs().max().item() <= 1e-3, ("mode 1", cin, cout, tune)
        n += 1
    for name, shape in (("MewZoom-2X-Ctrl", (2, 3, 9, 140)), ("MewZoom-3X-Ctrl", (1, 3, 7, 131)), ("MewZoom-4X-Ctrl", (1, 3, 6, 260))):
        cfg = dict(MODEL_CONFIGS[name])
        cfg["num_encoder_layers"] = 2
        torch.manual_seed(1)
        for rs in ("float32",):
            m = MewZoom(**cfg, residual_stream=rs).to(dev).eval()
            x, c = torch.rand(shape, generator=g).to(dev), torch.rand(shape[0], 3, generator=g).to(dev)
            y = m.upscale(x, c)
            m._flags_extra = _native.FLAG_SIMT_CONV
            ys = m.upscale(x, c)
            assert (y - ys).abs().max().item() <= 2e-3, (name, rs)
            y8 = m.upscale((x * 255).to(torch.uint8), c)
            assert y8.dtype == torch.uint8
            n += 1
    torch.cuda.synchronize()
    print(f"sanitize_cases: {n} cases ok")


if __name__ == "__main__":
    main()
