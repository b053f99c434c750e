"""Golden values of the xyz learning-rate schedule, produced by the reference's own get_expon_lr_func
(utils/general.py:31-64) imported from /root/reference in the build container.  Run: python tests/golden/make_lr_golden.py"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from utils.general import get_expon_lr_func  # noqa: E402

cases = [dict(lr_init=0.00016, lr_final=0.0000016, lr_delay_mult=0.01, max_steps=2990),                    # arguments.py:20-23
         dict(lr_init=0.00016 * 3.7, lr_final=0.0000016 * 3.7, lr_delay_mult=0.01, max_steps=2990),        # spatial_lr_scale
         dict(lr_init=0.01, lr_final=0.0001, lr_delay_steps=200, lr_delay_mult=0.05, max_steps=1000),
         dict(lr_init=0.0, lr_final=0.0, max_steps=10)]
steps = [-1, 0, 1, 7, 100, 199, 200, 999, 1000, 2989, 2990, 5000]
out = []
for c in cases:
    f = get_expon_lr_func(**c)
    out.append(dict(args=c, steps=steps, lr=[float(f(s)) for s in steps]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lr_golden.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", len(out), "cases")
