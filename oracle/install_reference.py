#!/usr/bin/env python
"""
TEST / MEASUREMENT INFRASTRUCTURE (not part of the product): installs the UNMODIFIED reference package
`shgyield` from the read-only checkout into the git-ignored directory baseline/_ref, so that
`bench.py --impl reference` (and the `cpu_baseline` leg) can time the reference's own code on the GPU box,
where /root/reference does not exist (baseline/_ref is git-ignored but travels with the gpurun snapshot).

    python oracle/install_reference.py [/root/reference]

The reference ships no setup.py / pyproject.toml, so `pip install /root/reference` fails ("neither 'setup.py' nor
'pyproject.toml' found").  The package directory is therefore copied to a scratch directory under /tmp, a
four-line setup.py is written THERE, and pip installs from that copy:

    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /tmp/<scratch>

Nothing of the reference enters the repository history.
"""
import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(REPO, "baseline", "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "shgyield", "shg.py"))


def install(src: str = "/root/reference", quiet: bool = True) -> bool:
    """Returns True when baseline/_ref/shgyield/shg.py exists afterwards."""
    pkg = os.path.join(src, "shgyield")
    if not os.path.isfile(os.path.join(pkg, "shg.py")):
        return installed()
    if installed() and open(os.path.join(pkg, "shg.py"), "rb").read() == open(os.path.join(TARGET, "shgyield", "shg.py"), "rb").read():
        return True
    scratch = tempfile.mkdtemp(prefix="sshg_reference_")
    try:
        shutil.copytree(pkg, os.path.join(scratch, "shgyield"), ignore=shutil.ignore_patterns("__pycache__"))
        with open(os.path.join(scratch, "setup.py"), "w") as f:
            f.write("from setuptools import setup\n"
                    "setup(name='shgyield-reference', version='0', packages=['shgyield'])\n")
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, scratch]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            if not quiet:
                sys.stderr.write(res.stdout + res.stderr)
            return installed()
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", quiet=False)
    print("baseline/_ref:", "installed" if ok else "NOT installed")
    sys.exit(0 if ok else 1)
