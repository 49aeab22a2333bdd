#!/bin/bash
# compute-sanitizer over the kernels at small sizes (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory
# hazards between the lanes / warps of a CTA; synccheck: barrier misuse).  Run on the GPU box: bash tools/sanitize.sh
set -u
cat > /tmp/sanitize_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
from oracle import lob_oracle
import __graft_entry__ as g
g.smoke()
oracle = lob_oracle.load()
for name, mac, dk in (("deep window", H.load_mac("hetero_deep_book"), dict(n_events=20000)),
                      ("deep 2 passes", H.load_mac("2_player_fq_fqc", nOrders=200, nTrades=64), dict(seed=9, n_events=20000, stress=True))):
    ld = H.load_for(mac, H.small_day(**dk))
    ref = H.OracleEnv(oracle, mac, ld, 40); gpu = H.CudaEnv(mac, ld, 40, ref.params)
    rng = np.random.default_rng(1)
    H.draw_prng(rng, ref.cfg, ref.arrays); gpu.set_inputs(ref.arrays); ref.reset(); gpu.reset()
    for s in range(66):
        H.draw_prng(rng, ref.cfg, ref.arrays); H.draw_actions(rng, ref.cfg, ref.arrays); gpu.set_inputs(ref.arrays)
        ref.step(n_threads=8); gpu.step()
    H.assert_arrays_match(ref.arrays, gpu.numpy(), ref.cfg, float_exact=True)
    print("ok", name, "second-pass environments so far:", int(gpu.numpy()["work_redo_count"][1]))
from jaxmarl_hft_b200 import lobster
m, rows, tm = lobster.preprocess_day_cuda(H.small_day(n_events=8000), device="cuda:0")
print("ok loader", tuple(m.shape))
PY
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/sanitize_case.py > /tmp/sanitize_$tool.log 2>&1
  echo "exit=$?"
  grep -E "^ok|^smoke|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error|Traceback" /tmp/sanitize_$tool.log | head -20
  tail -3 /tmp/sanitize_$tool.log
done
