"""Code-generation properties the measured speed rests on, checked on the built objects with cuobjdump
(no GPU needed): ptxas has to PROVE the replay kernel and the piped step's scan kernel warp-uniform -- the warp index goes
through a lane-0 shuffle for that (lob_book.cuh: uni()) -- or every REDUX gets a BRA.DIV + a software fallback, every branch a
BSSY / BSYNC pair and the kernels 10 more registers (28 warps per SM no longer fit)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "jaxmarl-hft_b200", "csrc", "build", "lob_inst_s4.o")


def _functions():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(OBJ):
        import __graft_entry__ as g
        g.build()
    if not (os.path.exists(OBJ) and os.path.exists(tool)):
        pytest.skip("needs the built csrc/build/lob_inst_s4.o and cuobjdump")
    text = subprocess.run([tool, "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    out = {}
    for chunk in text.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z0-9_.]+)", chunk)
        out[name] = ops
    return out


@pytest.mark.parametrize("kernel", ["lob_replay_kernelILi4E", "lob_step_scan_kernelILi4E"])
def test_scan_kernels_are_provably_warp_uniform(kernel):
    fns = _functions()
    names = [n for n in fns if kernel in n]
    assert names, f"{kernel} not found in {OBJ}"
    for n in names:
        ops = fns[n]
        assert ops.count("BRA.DIV") == 0, f"{n}: {ops.count('BRA.DIV')} BRA.DIV (ptxas no longer proves the scan warp-uniform)"
        assert ops.count("WARPSYNC.COLLECTIVE") == 0, n
        assert not any(o.startswith(("LDL", "STL")) for o in ops), f"{n}: local-memory traffic (spills) in the scan kernel"


def test_no_tensor_core_or_library_code_on_the_path():
    """north_star: no tensor cores (nothing here is a contraction): the kernels must not contain MMA instructions."""
    for n, ops in _functions().items():
        assert not any(o.startswith(("HMMA", "IMMA", "DMMA", "UTCMMA", "UTCHMMA", "QGMMA", "HGMMA")) for o in ops), n


def test_state_moves_with_the_bulk_copy_engine():
    """Books, trade logs' rows and message slices are staged by cp.async.bulk (SASS: UBLKCP) with completion on an mbarrier
    (SYNCS), not by per-thread loads: the replay kernel, the piped step's prep and scan kernels, the fused step kernel."""
    fns = _functions()
    for kernel in ("lob_replay_kernelILi4E", "lob_step_scan_kernelILi4E", "lob_step_prep_kernelILi4E", "lob_step_kernelILi4ELb0E"):
        names = [n for n in fns if kernel in n]
        assert names, kernel
        for n in names:
            assert any(o.startswith("UBLKCP") for o in fns[n]), f"{n}: no bulk copy"
            assert any(o.startswith("SYNCS") for o in fns[n]), f"{n}: no mbarrier"
