#!/usr/bin/env python
"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with
gpurun) for bench.py's reference arm and CPU baseline:

    python baseline/install_reference.py            # no-op when baseline/_ref is already there

The reference's build backend is poetry-core, which this image does not have: the package is installed
from a /tmp copy whose pyproject.toml build-system table is switched to setuptools (package sources
untouched: the script checks the installed tree against /root/reference file by file), with
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref`.  Nothing is copied into
the repository's history."""
import filecmp
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
PKG = "high_order_layers_torch"


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, PKG, "PolynomialLayers.py"))


def install(verbose: bool = True) -> bool:
    if installed():
        return True
    if not os.path.isdir(os.path.join(REF, PKG)):
        if verbose:
            print(f"install_reference: {REF} is not here; bench.py falls back to the numpy oracle port", file=sys.stderr)
        return False
    tmp = "/tmp/hol_refcopy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    pp = os.path.join(tmp, "pyproject.toml")
    text = open(pp).read()
    text = text.replace('requires = ["poetry-core>=1.0.0"]', 'requires = ["setuptools"]')
    text = text.replace('build-backend = "poetry.core.masonry.api"', 'build-backend = "setuptools.build_meta"')
    text += ('\n[project]\nname = "high-order-layers-torch"\nversion = "2.6.0"\n\n[tool.setuptools]\n'
             f'packages = ["{PKG}", "{PKG}.sparse_optimizers"]\n')
    open(pp, "w").write(text)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--target", DST, tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    if r.returncode != 0 or not installed():
        if verbose:
            print("install_reference: pip failed:\n" + r.stdout[-2000:] + r.stderr[-2000:], file=sys.stderr)
        return False
    cmp = filecmp.dircmp(os.path.join(DST, PKG), os.path.join(REF, PKG), ignore=["__pycache__"])
    clean = not (cmp.diff_files or cmp.left_only or cmp.right_only)
    if verbose:
        print(f"install_reference: installed into {DST}; identical to {REF}/{PKG}: {clean}")
    return clean


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
