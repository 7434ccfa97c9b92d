import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_cases as pc
(g, d, dg, sub_g), (og, od, ocomp, osub), z, sX, sy, ny = pc.burst_case(64, 8)
for name, (pm, om, x, y) in [('D', (d, od, sX, sy)), ('subG', (sub_g, osub, z, ny)), ('DG', (dg, ocomp, z, [1]*8))]:
    try:
        errs, w0 = pc.compare_step(pm, om, x, y, check_predict=False)
        print(name, 'OK', {k: '%.1e' % v for k, v in errs.items()})
    except AssertionError as e:
        print(name, 'FAIL', str(e)[:1500])
