"""Label + pairs on a random grid with very many labels (development aid): python tools/many_labels.py [n] [p]."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.35
torch.manual_seed(1)
bits = dev.BitGrid.from_bool(torch.rand((n, n, n), device="cuda") < p)
plan, hdr = dev.label_with_retry(bits)
for name, fn in (("label", lambda: plan.label(bits)), ("pairs", lambda: plan.pairs(bits))):
    torch.cuda.synchronize()
    t = time.time()
    fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.time() - t) * 1e3:.2f} ms")
print("runs", hdr.total_runs, "labels", hdr.n_true, hdr.n_false, "contacts", hdr.n_contacts, "bridges", hdr.n_bridges,
      "caps", plan.caps.contact_slots, plan.caps.bridge_slots)
c = plan.contacts()
k = c[:, 0].astype(np.int64) * (1 << 32) + (c[:, 1].astype(np.int64) + (1 << 31))
print("contacts sorted and unique:", bool(np.all(np.diff(k) > 0)), len(c))
