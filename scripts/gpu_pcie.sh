#!/bin/bash
# PCIe ceiling of the box: pinned -> device copies, one stream vs two, large vs ring-sized pieces.
python - <<'PY'
import torch, time
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, pin_memory=True); src.fill_(1)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(piece, streams, reps=6):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        for k, off in enumerate(range(0, n, piece)):
            with torch.cuda.stream(ss[k % streams]):
                dst[off:off + piece].copy_(src[off:off + piece], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n / dt / 1e9
for piece in (1 << 30, 64 << 20, 16 << 20, 4 << 20, 1 << 20):
    print(f"piece {piece >> 20:5d} MiB: 1 stream {run(piece, 1):.1f} GB/s, 2 streams {run(piece, 2):.1f} GB/s, 4 streams {run(piece, 4):.1f} GB/s")
PY
