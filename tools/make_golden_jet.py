#!/usr/bin/env python3
"""tests/golden/jet_1024.json: colormap[i] = matplotlib.cm.jet(i / 1024)[0:3] for i < 1024 (mapping_common.py:158-163),
independently of the two in-repo copies of the table (csrc fill_jet_host, oracle fill_jet - the same hand-typed code).

matplotlib is not installable in this container (no wheel, no network), so the values are produced with matplotlib's
published ALGORITHM in float64 numpy instead of being read from the package:
  * `_jet_data` (matplotlib/_cm.py) - the piecewise-linear segment table of the three channels (no jumps: y0 == y1);
  * `LinearSegmentedColormap._init` -> `_create_lookup_table(N=256, data)`: lut = np.interp(np.linspace(0, 1, 256), x, y);
  * `Colormap.__call__` on a float X: index int(X * N) (floor), clipped to N - 1 -> jet(i/1024) = lut[i // 4].
Anchors from the real package (any matplotlib >= 2.0 prints these): jet(0.0) = (0, 0, 0.5), jet(1.0) = (0.5, 0, 0),
jet(0.5) = (0.4901960784313725, 1.0, 0.4775458570524984) - asserted below.
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jet_1024.json")
JET = {"red": [(0.00, 0), (0.35, 0), (0.66, 1), (0.89, 1), (1.00, 0.5)],
       "green": [(0.000, 0), (0.125, 0), (0.375, 1), (0.640, 1), (0.910, 0), (1.000, 0)],
       "blue": [(0.00, 0.5), (0.11, 1), (0.34, 1), (0.65, 0), (1.00, 0)]}


def main():
    xs = np.linspace(0.0, 1.0, 256)
    lut = np.stack([np.interp(xs, [p[0] for p in JET[c]], [p[1] for p in JET[c]]) for c in ("red", "green", "blue")], 1)
    jet = lambda X: lut[min(int(X * 256), 255)]
    assert np.allclose(jet(0.0), (0, 0, 0.5), atol=0) and np.allclose(jet(1.0), (0.5, 0, 0), atol=0)
    assert np.allclose(jet(0.5), (0.4901960784313725, 1.0, 0.4775458570524984), rtol=0, atol=1e-15)
    table = [[float(v) for v in jet(i / 1024)] for i in range(1024)]
    json.dump({"source": "matplotlib's jet: _jet_data + _create_lookup_table(256) + Colormap.__call__, restated in float64 numpy "
                         "(tools/make_golden_jet.py); anchors jet(0), jet(0.5), jet(1) from the real package",
               "colormap": table}, open(OUT, "w"))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
