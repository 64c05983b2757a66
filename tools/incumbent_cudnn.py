"""The GPU incumbent beside our kernels: the same flat architecture written in plain torch.nn and run
eagerly on the B200 through cuDNN -- fp32 with TF32 off and on, and bf16 / fp16 channels_last.  A measurement tool: it
imports neither the product path nor oracle/; the module below is the architecture of SURVEY.md Appendix C
(bicubic + stem 1x1 + L x [conv3x3 -> FiLM -> SiLU -> conv3x3 -> + x] + conv3x3 -> PixelShuffle, + skip, clamp).

    python tools/incumbent_cudnn.py [cfg2|cfg4a|cfg3|tiny] [steps]

Prints one JSON line per variant: ms per step (CUDA events, 3 warm-ups) and output Mpx/s.
"""
import json
import sys

import torch
from torch import nn

WORKLOADS = {  # name: (ratio, channels, layers, batch, H, W)
    "cfg2": (2, 48, 20, 16, 540, 960),
    "cfg3": (3, 54, 30, 1, 720, 1280),
    "cfg4a": (4, 96, 40, 1, 540, 960),
    "tiny": (2, 48, 20, 1, 256, 256),
}


class Block(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, 2 * ch, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(2 * ch, ch, 3, padding=1, bias=False)
        self.film = nn.Linear(3, 4 * ch)
        self.act = nn.SiLU()

    def forward(self, x, c):
        gamma, beta = self.film(c).to(x.dtype).chunk(2, dim=1)
        h = self.conv1(x) * (1 + gamma[:, :, None, None]) + beta[:, :, None, None]
        return x + self.conv2(self.act(h))


class Flat(nn.Module):
    def __init__(self, r: int, ch: int, layers: int):
        super().__init__()
        self.up = nn.Upsample(scale_factor=r, mode="bicubic")
        self.stem = nn.Conv2d(3, ch, 1)
        self.blocks = nn.ModuleList(Block(ch) for _ in range(layers))
        self.head = nn.Conv2d(ch, 3 * r * r, 3, padding=1, bias=False)
        self.shuffle = nn.PixelShuffle(r)

    @torch.inference_mode()
    def upscale(self, x, c, dtype):
        s = self.up(x)
        z = self.stem(x.to(dtype))
        for b in self.blocks:
            z = b(z, c)
        return torch.clamp(s + self.shuffle(self.head(z)).float(), 0, 1)


def main() -> None:
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    r, ch, layers, B, H, W = WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = torch.rand(B, 3, H, W, device=dev)
    c = torch.tensor([[0.5, 0.2, 0.3]], device=dev).repeat(B, 1)
    variants = [("fp32 (TF32 off)", torch.float32, False, False), ("fp32 (TF32 on)", torch.float32, True, False),
                ("bf16 channels_last", torch.bfloat16, True, True), ("fp16 channels_last", torch.float16, True, True)]
    torch.backends.cudnn.benchmark = True
    for label, dtype, tf32, cl in variants:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        m = Flat(r, ch, layers).to(dev).eval()
        m.blocks.to(dtype)
        m.stem.to(dtype)
        m.head.to(dtype)
        for b in m.blocks:
            b.film.float()
        if cl:
            m = m.to(memory_format=torch.channels_last)
        xi = x.contiguous(memory_format=torch.channels_last) if cl else x
        try:
            for _ in range(3):
                m.upscale(xi, c, dtype)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                m.upscale(xi, c, dtype)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            print(json.dumps({"incumbent": "torch %s eager / cuDNN %s" % (torch.__version__, torch.backends.cudnn.version()),
                              "variant": label, "workload": wl, "ms_per_step": round(ms, 3),
                              "output_mpx_per_s": round(B * H * W * r * r / ms / 1e3, 1)}), flush=True)
        except Exception as e:  # an out-of-memory variant must not take the others with it
            print(json.dumps({"variant": label, "workload": wl, "error": str(e)[:200]}), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
