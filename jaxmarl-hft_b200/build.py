"""In-tree build of the CUDA C-ABI library ``csrc/liblobstep.so`` (sm_100a only).

``nvcc`` cross-compiles here without a GPU; the built ``.so`` is git-ignored but travels to the GPU box with the
snapshot.  One object per book-capacity class (LOB_SLOTS = rows per lane) so the instantiations build in parallel.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SO = os.path.join(CSRC, "liblobstep.so")
OBJ = os.path.join(CSRC, "build")
SLOTS = (1, 2, 4, 8, 16)
GROUPED = ((8, 14),)   # (lanes per book, rows per lane) classes of the grouped kernels
# --fmad=false: a*b+c stays two roundings, as in the XLA lowering of the reference's float32 expressions
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("LOB_NVCC_EXTRA", "").split()


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
        os.path.join(INCLUDE, "lobstep.h")]


def _digest():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in _sources():
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_variant(name, defines=(), slots=SLOTS):
    """A development build with extra ``-D`` flags -> csrc/variants/liblobstep_<name>.so (git-ignored; travels with gpurun).
    Use it with ``LOB_SO=<path> python tools/kbench.py`` to A/B a kernel change against the committed build."""
    vdir = os.path.join(CSRC, "variants")
    return build_cuda(force=True, so=os.path.join(vdir, f"liblobstep_{name}.so"), obj=os.path.join(vdir, f"obj_{name}"),
                      extra=[f"-D{d}" for d in defines], slots=slots)


def build_cuda(force=False, verbose=False, so=SO, obj=OBJ, extra=(), slots=SLOTS):
    """Compile csrc/*.cu -> csrc/liblobstep.so.  Skipped when the sources are unchanged since the last build."""
    SO, OBJ = so, obj   # noqa: N806  (the variant build reuses the body below with its own paths)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest() + " ".join(extra)
    if not force and os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read() == digest:
        return SO
    os.makedirs(OBJ, exist_ok=True)
    extra = list(extra)
    jobs = [(os.path.join(CSRC, "lobstep.cu"), os.path.join(OBJ, "lobstep.o"), extra),
            (os.path.join(CSRC, "lob_loader.cu"), os.path.join(OBJ, "lob_loader.o"), [])]
    for s in SLOTS:   # (every capacity class is linked; `slots` only says which ones get the extra flags)
        jobs.append((os.path.join(CSRC, "lob_inst.cu"), os.path.join(OBJ, f"lob_inst_s{s}.o"),
                     [f"-DLOB_SLOTS={s}"] + (extra if s in slots else [])))

    for l, r in GROUPED:
        jobs.append((os.path.join(CSRC, "lob_ginst.cu"), os.path.join(OBJ, f"lob_ginst_l{l}r{r}.o"), [f"-DLOB_GL={l}", f"-DLOB_GR={r}"]))

    def compile_one(job):
        src, obj, extra = job
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for obj, log in ex.map(compile_one, jobs):
            logs.append(f"== {os.path.basename(obj)}\n{log}")
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    cmd = [_nvcc(), "-shared", "-o", SO] + [j[1] for j in jobs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(logs))
    return SO


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
