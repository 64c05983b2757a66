"""GPU bring-up diagnostics: run on the B200 box, one group per process so that a faulting kernel
cannot poison the others.

    python tools/gpu_diag.py all            # spawns every group below with its own timeout
    python tools/gpu_diag.py probe|rate|small|simt|tc <halo_mode>|e2e <halo_mode|simt>

Prints one line per check:  [group] name ... value  OK|FAIL
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _report(group, name, value, ok):
    print(f"[{group}] {name:<58s} {value}  {'OK' if ok else 'FAIL'}", flush=True)


def g_probe():
    from ultrazoom_b200 import ops

    for kc in (64, 32, 16):
        for shift in (0, 1, 2, 3, 5, 8, 130, 131, 132, 256):
            for bo in (0, 1):
                try:
                    err = ops.probe_umma(kc, shift, bo)
                except Exception as e:  # noqa: BLE001
                    _report("probe", f"kc={kc} shift={shift} bo_mode={bo}", f"EXC {e}", False)
                    return
                _report("probe", f"kc={kc} shift={shift} bo_mode={bo}", f"max_err={err:.4g}", err == 0.0)


def g_rate():
    from ultrazoom_b200 import ops

    for n in (48, 96, 128, 192, 256):
        for kc in (64, 32, 16):
            for da, dd in ((1, 1), (2, 1), (1, 2), (2, 2)):
                if n * dd > 512:
                    continue
                cyc = ops.probe_mma_rate(n, kc, 4000, 148, da, dd)
                _report("rate", f"n={n} kc={kc} distinct_a={da} distinct_d={dd}", f"cyc/mma={cyc:.1f} (N/2={n / 2:.0f})", True)


def g_gap():
    """How long may the issuing thread stay away between two bursts before the tensor pipe runs dry?"""
    from ultrazoom_b200 import _native, ops

    lib = _native.load()
    for n in (48, 96, 192):
        for burst in (4, 8):                      # 32 / 64 UMMAs per burst
            for commit in (0, 1):
                row = []
                for gap in (0, 50, 100, 150, 200, 300, 400, 600):
                    lib.mz_probe_set_gap(burst, gap, commit)
                    cyc = ops.probe_mma_rate(n, 32, 4000, 148, 2, 2, 0)
                    row.append(f"{gap}:{cyc * 8 * burst:.0f}")
                _report("gap", f"n={n} burst={8 * burst} commit={commit}", "cycles per burst by gap  " + " ".join(row), True)
    lib.mz_probe_set_gap(0, 0, 0)


def g_shift():
    """UMMA rate when the A descriptor starts a few rows into the swizzled tile (the shared-halo taps)."""
    from ultrazoom_b200 import ops

    for n in (48, 96, 192):
        for kc in (64, 32, 16):
            row = " ".join(f"{s}:{ops.probe_mma_rate(n, kc, 4000, 148, 1, 2, s):.1f}" for s in (0, 1, 2, 3, 4, 8, 10, 16))
            _report("shift", f"n={n} kc={kc}", f"cyc/mma by a_row_shift  {row}", True)


def _rand_bf16(shape, gen, scale=1.0):
    import torch

    return (torch.randn(shape, generator=gen) * scale).to(torch.bfloat16)


def g_small():
    import torch
    from torch.nn import functional as F

    from ultrazoom_b200 import ops

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    for r in (2, 3, 4):
        for shape in ((1, 3, 5, 7), (2, 3, 33, 61), (1, 3, 128, 256)):
            x = torch.rand(shape, generator=g)
            ref = F.interpolate(x, scale_factor=r, mode="bicubic")
            got = ops.bicubic(x.to(dev), r).cpu()
            err = (got - ref).abs().max().item()
            _report("small", f"bicubic r={r} shape={shape}", f"max_err={err:.3g}", err <= 2e-6)
    for Cc in (16, 48, 54, 96):
        x = torch.rand(2, 3, 9, 11, generator=g)
        w = torch.randn(Cc, 3, 1, 1, generator=g) * 0.5
        b = torch.randn(Cc, generator=g) * 0.1
        ref = F.conv2d(x, w, b).permute(0, 2, 3, 1)
        zf, zb = ops.stem_pack(x.to(dev), w, b)
        err = (zf.cpu()[..., :Cc] - ref).abs().max().item()
        errb = (zb.cpu().float()[..., :Cc] - ref).abs().max().item()
        padz = zf.cpu()[..., Cc:].abs().max().item() if zf.shape[-1] > Cc else 0.0
        _report("small", f"stem C={Cc}", f"fp32_err={err:.3g} bf16_err={errb:.3g} pad={padz}", err <= 1e-5 and errb <= 2e-2 and padz == 0)
    L, B, Fc, hC = 3, 2, 3, 108
    w = torch.randn(L, 2 * hC, Fc, generator=g)
    bb = torch.randn(L, 2 * hC, generator=g)
    c = torch.rand(B, Fc, generator=g)
    film = ops.control_film(c.to(dev), w.to(dev), bb.to(dev), B).cpu()
    gl = torch.einsum("bf,lnf->lbn", c, w) + bb[:, None, :]
    err = max((film[:, :, 0, :hC] - (1 + gl[..., :hC])).abs().max().item(), (film[:, :, 1, :hC] - gl[..., hC:]).abs().max().item())
    pad_ok = bool((film[:, :, 0, hC:] == 1).all() and (film[:, :, 1, hC:] == 0).all())
    _report("small", "film L=3 B=2 hC=108", f"max_err={err:.3g} pad_ok={pad_ok}", err <= 1e-5 and pad_ok)


def _conv_ref(inp_bf16, w, mode, film, zf0):
    """CPU fp32 reference on the SAME bf16-rounded operands."""
    import torch
    from torch.nn import functional as F

    x = inp_bf16.float().permute(0, 3, 1, 2)[:, :w.shape[1]]
    acc = F.conv2d(x, w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    if mode == 0:
        if film is not None:
            acc = acc * film[:, 0][:, None, None, :acc.shape[-1]] + film[:, 1][:, None, None, :acc.shape[-1]]
        return F.silu(acc), None
    z = zf0[..., :acc.shape[-1]] + acc
    return z, z


def _conv_case(group, use_tc, tune_kw, cin, cout, mode, shape, gen, with_film=True):
    import torch

    from ultrazoom_b200 import _native, ops

    dev = torch.device("cuda", 0)
    B, H, W = shape
    cin_p, cout_p = ops.padded_channels(cin), ops.padded_channels(cout)
    inp = torch.zeros(B, H, W, cin_p, dtype=torch.bfloat16)
    inp[..., :cin] = _rand_bf16((B, H, W, cin), gen)
    w = torch.randn(cout, cin, 3, 3, generator=gen) / (3.0 * cin ** 0.5)
    wp = ops.pack_conv_weight(w, dev, dtype=torch.bfloat16)
    film = None
    if mode == 0 and with_film:
        film = torch.ones(B, 2, cout_p)
        film[:, 0, :cout] = 1 + 0.3 * torch.randn(B, cout, generator=gen)
        film[:, 1, :] = 0
        film[:, 1, :cout] = 0.3 * torch.randn(B, cout, generator=gen)
    zf0 = torch.zeros(B, H, W, cout_p)
    zf0[..., :cout] = torch.randn(B, H, W, cout, generator=gen)
    ref, zref = _conv_ref(inp, w, mode, film, zf0)
    tune = _native.tune(**tune_kw) if tune_kw is not None else None
    zf = zf0.to(dev).contiguous()
    out = ops.conv3x3(inp.to(dev), wp, mode, film.to(dev) if film is not None else None, zf if mode == 1 else None,
                      use_tc=use_tc, tune=tune)
    torch.cuda.synchronize()
    got = out.cpu().float()
    err = (got[..., :cout] - ref).abs().max().item()
    padv = got[..., cout:].abs().max().item() if cout_p > cout else 0.0
    ok = err <= 3e-2 and padv == 0.0
    extra = ""
    if mode == 1:
        e2 = (zf.cpu()[..., :cout] - zref).abs().max().item()
        extra = f" zf_err={e2:.3g}"
        ok = ok and e2 <= 2e-4
    _report(group, f"conv cin={cin} cout={cout} mode={mode} shape={shape} tune={tune_kw}", f"max_err={err:.3g} pad={padv}{extra}", ok)
    return ok


def _head_case(group, use_tc, tune_kw, cin, r, shape, gen, skip_mode, clamp):
    import torch
    from torch.nn import functional as F

    from ultrazoom_b200 import _native, ops

    dev = torch.device("cuda", 0)
    B, H, W = shape
    cin_p = ops.padded_channels(cin)
    zb = torch.zeros(B, H, W, cin_p, dtype=torch.bfloat16)
    zb[..., :cin] = _rand_bf16((B, H, W, cin), gen)
    w = torch.randn(3 * r * r, cin, 3, 3, generator=gen) / (3.0 * cin ** 0.5)
    x = torch.rand(B, 3, H, W, generator=gen)
    wp = ops.pack_conv_weight(w, dev, dtype=torch.bfloat16)
    u = F.pixel_shuffle(F.conv2d(zb.float().permute(0, 3, 1, 2)[:, :cin], w.to(torch.bfloat16).float(), padding=1), r)
    s = F.interpolate(x, scale_factor=r, mode="bicubic")
    ref = u + (s if skip_mode else 0)
    if clamp:
        ref = ref.clamp(0, 1)
    tune = _native.tune(**tune_kw) if tune_kw is not None else None
    y = None
    if skip_mode == 1:
        y = ops.bicubic(x.to(dev), r)
    got = ops.head_shuffle_add(zb.to(dev), wp, r, x=x.to(dev), y=y, skip_mode=skip_mode, clamp01=clamp, use_tc=use_tc,
                               tune=tune)
    torch.cuda.synchronize()
    err = (got.cpu() - ref).abs().max().item()
    ok = err <= 2e-4
    _report(group, f"head cin={cin} r={r} shape={shape} skip={skip_mode} clamp={clamp} tune={tune_kw}", f"max_err={err:.3g}", ok)
    return ok


def g_simt():
    import torch

    g = torch.Generator().manual_seed(2)
    _conv_case("simt", False, None, 16, 32, 0, (1, 6, 9), g)
    _conv_case("simt", False, None, 48, 96, 0, (2, 13, 150), g)
    _conv_case("simt", False, None, 108, 54, 1, (1, 9, 40), g)
    _conv_case("simt", False, None, 192, 96, 1, (1, 7, 33), g)
    for r in (2, 3, 4):
        _head_case("simt", False, None, 48, r, (2, 7, 19), g, 2, True)
        _head_case("simt", False, None, 54, r, (1, 5, 9), g, 1, False)
        _head_case("simt", False, None, 16, r, (1, 5, 9), g, 0, False)


def g_tc(halo):
    import torch

    g = torch.Generator().manual_seed(3)
    grp = f"tc{halo}"
    # smallest possible first: one patch, one chunk
    _conv_case(grp, True, dict(halo_mode=halo, rows=1, acc_stages=1), 16, 16, 0, (1, 1, 128), g, with_film=False)
    _conv_case(grp, True, dict(halo_mode=halo, rows=1, acc_stages=1), 16, 16, 0, (1, 3, 128), g, with_film=False)
    _conv_case(grp, True, dict(halo_mode=halo, rows=1), 32, 32, 0, (1, 5, 130), g)
    _conv_case(grp, True, dict(halo_mode=halo, rows=1), 64, 64, 0, (1, 5, 130), g)
    _conv_case(grp, True, dict(halo_mode=halo, rows=2), 64, 64, 0, (1, 5, 130), g)
    _conv_case(grp, True, dict(halo_mode=halo), 48, 96, 0, (2, 13, 150), g)
    _conv_case(grp, True, dict(halo_mode=halo), 96, 48, 1, (2, 13, 150), g)
    _conv_case(grp, True, dict(halo_mode=halo), 96, 192, 0, (1, 21, 300), g)
    _conv_case(grp, True, dict(halo_mode=halo), 192, 96, 1, (1, 21, 300), g)
    _conv_case(grp, True, dict(halo_mode=halo), 54, 108, 0, (1, 10, 70), g)
    _conv_case(grp, True, dict(halo_mode=halo), 108, 54, 1, (1, 10, 70), g)
    _conv_case(grp, True, dict(halo_mode=halo, max_ctas=3), 96, 192, 0, (2, 30, 260), g)   # many patches per CTA
    _conv_case(grp, True, dict(halo_mode=halo, max_ctas=3), 192, 96, 1, (2, 30, 260), g)
    for r in (2, 3, 4):
        _head_case(grp, True, dict(halo_mode=halo), 48, r, (2, 7, 190), g, 2, True)
        _head_case(grp, True, dict(halo_mode=halo), 96, r, (1, 9, 129), g, 1, False)


def g_e2e(which):
    import torch

    from oracle import max_abs_err, psnr
    from tests.helpers import CASES, load_case, oracle_from_case
    from ultrazoom_b200 import MewZoom, _native

    dev = torch.device("cuda", 0)
    for name in CASES:
        cfg, sd, x, c, out = load_case(name)
        m = MewZoom(**cfg)
        m.load_state_dict(sd)
        m = m.to(dev)
        if which == "simt":
            m._flags_extra = _native.FLAG_SIMT_CONV
        else:
            m.set_conv_tune(-1, dev, halo_mode=int(which))
        y = m.forward(x.to(dev), c.to(dev) if c is not None else None).cpu()
        yc = m.upscale(x.to(dev), c.to(dev) if c is not None else None).cpu()
        e1, e2 = max_abs_err(y, out["forward"]), max_abs_err(yc, out["upscale"])
        _report(f"e2e-{which}", f"{name}", f"forward_err={e1:.3g} upscale_err={e2:.3g} psnr={psnr(yc, out['upscale']):.1f}",
                e1 <= 2e-2 and e2 <= 2e-2)
    # one deeper model vs the oracle
    from oracle import make_oracle

    cfgd = dict(upscale_ratio=4, num_channels=96, hidden_ratio=2, num_encoder_layers=8, control_features=3)
    o = make_oracle(cfgd, seed=0)
    m = MewZoom(**cfgd)
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    if which == "simt":
        m._flags_extra = _native.FLAG_SIMT_CONV
    else:
        m.set_conv_tune(-1, dev, halo_mode=int(which))
    x = torch.rand(1, 3, 40, 150, generator=torch.Generator().manual_seed(5))
    c = torch.tensor([0.5, 0.2, 0.3])
    ref = o.upscale(x, c)
    t0 = time.time()
    got = m.upscale(x.to(dev), c.to(dev)).cpu()
    e = max_abs_err(got, ref)
    _report(f"e2e-{which}", "4X-Ctrl 96ch L=8 40x150", f"err={e:.3g} psnr={psnr(got, ref):.1f} ({time.time() - t0:.2f}s)", e <= 2e-2)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        groups = [["small"], ["simt"], ["e2e", "simt"], ["tc", "0"], ["e2e", "0"], ["tc", "1"], ["e2e", "1"]]
        for gsel in groups:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), *gsel], timeout=240, capture_output=True,
                                   text=True)
                sys.stdout.write(r.stdout)
                tail = "\n".join(r.stderr.strip().splitlines()[-8:])
                print(f"== group {' '.join(gsel)}: exit {r.returncode} in {time.time() - t0:.1f}s" + (f"\n{tail}" if r.returncode else ""), flush=True)
            except subprocess.TimeoutExpired as e:
                sys.stdout.write((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
                print(f"== group {' '.join(gsel)}: TIMEOUT", flush=True)
        return
    if what == "probe":
        g_probe()
    elif what == "rate":
        g_rate()
    elif what == "shift":
        g_shift()
    elif what == "gap":
        g_gap()
    elif what == "small":
        g_small()
    elif what == "simt":
        g_simt()
    elif what == "tc":
        g_tc(int(sys.argv[2]))
    elif what == "e2e":
        g_e2e(sys.argv[2])


if __name__ == "__main__":
    main()
