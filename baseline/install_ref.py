"""Install the UNMODIFIED reference (torch_specinv 0.2.1) into the git-ignored baseline/_ref/ so that it travels to
the GPU box with the working tree (there is no /root/reference there).

    python baseline/install_ref.py [--force]

Equivalent to
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
(`--no-deps`: torch and tqdm are in the image, there is no index to resolve them from; the install runs from a copy
under /tmp because /root/reference is read-only and setuptools writes build/ and *.egg-info into the source tree).
Nothing of the reference's sources is copied into the tracked tree.  bench.py --impl reference and the
`reference_cuda` leg import it from there; without it they fall back to the oracle port and say so."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SRC = os.environ.get("SPECINV_REFERENCE_SRC", "/root/reference")


def installed() -> bool:
    return os.path.exists(os.path.join(TARGET, "torch_specinv", "methods.py"))


def install(force: bool = False) -> bool:
    """True when baseline/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(SRC):
        return False
    tmp = tempfile.mkdtemp(prefix="specinv_ref_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(SRC, work)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            return False
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else "NOT installed")
    sys.exit(0 if ok else 1)
