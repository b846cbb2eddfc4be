"""TEST INFRASTRUCTURE ONLY — compile the C restatements under oracle/ with gcc into oracle/_build/ (git-ignored; the
built .so travels to the GPU box with the snapshot).  -ffp-contract=off: the oracle's arithmetic contract is plain IEEE
fp32 in the written order (see raster_oracle.c)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
TARGETS = {"libraster_oracle.so": ["raster_oracle.c"]}


def lib_path(name="libraster_oracle.so"):
    return os.path.join(OUT_DIR, name)


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    gcc = shutil.which("gcc")
    for out, srcs in TARGETS.items():
        dst = os.path.join(OUT_DIR, out)
        srcs = [os.path.join(HERE, s) for s in srcs]
        if not force and os.path.exists(dst) and all(os.path.getmtime(s) <= os.path.getmtime(dst) for s in srcs):
            continue
        if gcc is None:
            raise RuntimeError("gcc not found; cannot build the oracle")
        cmd = [gcc, "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
               "-o", dst + ".tmp"] + srcs + ["-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
        os.replace(dst + ".tmp", dst)
    return OUT_DIR


if __name__ == "__main__":
    print(build(force=True))
