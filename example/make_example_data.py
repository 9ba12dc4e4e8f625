#!/usr/bin/env python
"""Writes the Si(111)(1x1):H example spectra as text files (the formats `loadeps` / `loadshg` read)
from example/si111_spectra.npz (the raw file columns of the example material), next to example/si111-input.yml:

    python example/make_example_data.py [target_dir]
    python shgyield.py example/si111-input.yml
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def write(target=HERE):
    z = np.load(os.path.join(HERE, "si111_spectra.npz"))
    e = z["energy"]
    for sub in ("chi1-linear", "chi2-nonlinear"):
        os.makedirs(os.path.join(target, sub), exist_ok=True)
    for name in z.files:
        if "-chi1-" in name and not name.startswith("test-"):
            np.savetxt(os.path.join(target, "chi1-linear", name), np.column_stack([e] + list(z[name])),
                       fmt="%.5f    %.6E    %.6E", header="energy1w    re    im")
        elif "-chi2-" in name and not name.startswith("test-"):
            np.savetxt(os.path.join(target, "chi2-nonlinear", name), np.column_stack([e] + list(z[name])),
                       fmt="%10.5f   %13.6E   %13.6E   %13.6E   %13.6E", header="energy1w  1w_re  1w_im  2w_re  2w_im")
    return target


if __name__ == "__main__":
    print(write(sys.argv[1] if len(sys.argv) > 1 else HERE))
