"""Where the wall-clock of a drop-in ``Containment(...)`` constructor goes (development aid):

    python tools/api_profile.py [n=1024]

cProfile of the third constructor call for both flag sets, with a cuda synchronize on both sides."""
import contextlib
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import Containment, synth  # noqa: E402
from mdvcontainment_b200.universe import Universe  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    for name, kw in (("no_mapping", dict(no_mapping=True, betafactors=False)), ("default flags", dict())):
        for rep in range(3):
            u = Universe(pos.copy(), dims)
            pr = cProfile.Profile()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if rep == 2:
                pr.enable()
            with contextlib.redirect_stdout(io.StringIO()):
                c = Containment(u.atoms, 0.5, closing=True, **kw)
                s = str(c)
            torch.cuda.synchronize()
            pr.disable()
            print(f"{name} [{rep}] {1e3 * (time.perf_counter() - t0):8.1f} ms")
            if rep == 2:
                out = io.StringIO()
                pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(28)
                print("\n".join(l for l in out.getvalue().splitlines() if l.strip())[:6000])
            del c, u
    print(s)


if __name__ == "__main__":
    main()
