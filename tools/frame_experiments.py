"""What-if timings of the frame loop (development aid, not a benchmark):

    python tools/frame_experiments.py [n=1024]

* frames in flight 1..5
* the same loop with the binning (clear + bucket + RED kernels) removed -- the bit grid of the previous frame is reused --
  which bounds what a better binning could buy inside the overlapped loop
* the same loop without the dense write, which shows the cost of everything else when nothing competes with it
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth  # noqa: E402
from mdvcontainment_b200.pipeline import FramePipeline  # noqa: E402


def loop(pipes, pos, steps):
    pending = []
    for i in range(steps):
        p = pipes[i % len(pipes)]
        if len(pending) == len(pipes):
            pending.pop(0).finish()
        p.begin(pos)
        pending.append(p)
    while pending:
        pending.pop(0).finish()


def timed(pipes, pos, steps=40):
    loop(pipes, pos, 4 * len(pipes))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop(pipes, pos, steps)
    for p in pipes:
        torch.cuda.current_stream().wait_stream(p.stream)
        torch.cuda.current_stream().wait_stream(p.stream_x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    pos_h, dims = synth.vesicle_scene(box_nm=n / 2.0)
    pos = torch.from_numpy(pos_h).cuda()
    for depth in (1, 2, 3, 4, 5):
        pipes = [FramePipeline(dims, 0.5, closing=True) for _ in range(depth)]
        print(f"in flight {depth}: {timed(pipes, pos):7.3f} ms/frame")
        del pipes
        torch.cuda.empty_cache()

    one = torch.cuda.Stream()
    for depth in (2, 3, 4):
        pipes = [FramePipeline(dims, 0.5, closing=True, streams=(one, one)) for _ in range(depth)]
        t = timed(pipes, pos)
        ex = sum(p.expand_elapsed_ms() for p in pipes) / depth
        print(f"{depth} in flight on ONE shared stream: {t:7.3f} ms/frame (dense write inside the loop {ex:6.3f} ms)")
        del pipes
        torch.cuda.empty_cache()
    two = (torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0))
    pipes = [FramePipeline(dims, 0.5, closing=True, streams=two) for _ in range(3)]
    t = timed(pipes, pos)
    ex = sum(p.expand_elapsed_ms() for p in pipes) / 3
    print(f"3 in flight on TWO shared streams (stage 1 / dense write): {t:7.3f} ms/frame (dense write {ex:6.3f} ms)")
    del pipes
    torch.cuda.empty_cache()

    pipes = [FramePipeline(dims, 0.5, closing=True) for _ in range(3)]
    base = timed(pipes, pos)
    print(f"(own streams: dense write inside the loop {sum(p.expand_elapsed_ms() for p in pipes) / 3:6.3f} ms)")
    # no binning: keep bits_a of the last frame (graphs are re-captured without the memset and the RED kernel)
    real_bin = dev.bin_grid
    for p in pipes:
        p._graphs.clear()
    dev.bin_grid = lambda *a, **k: (None, None)
    no_bin = timed(pipes, pos)
    dev.bin_grid = real_bin
    print(f"3 in flight: {base:7.3f} ms/frame; without binning {no_bin:7.3f} ms/frame ({base - no_bin:+.3f})")
    del pipes
    torch.cuda.empty_cache()

    pipes = [FramePipeline(dims, 0.5, closing=True) for _ in range(3)]
    for p in pipes:
        p.plan.expand = lambda *a, **k: None
    print(f"3 in flight, no dense write: {timed(pipes, pos):7.3f} ms/frame (stage 1 + host only)")
    for p in pipes:
        p.labels, p.write_labels = None, False
    del pipes
    torch.cuda.empty_cache()
    pipes = [FramePipeline(dims, 0.5, closing=True, write_labels=False) for _ in range(3)]
    print(f"3 in flight, components only: {timed(pipes, pos):7.3f} ms/frame")


if __name__ == "__main__":
    main()
