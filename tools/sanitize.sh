#!/bin/bash
# compute-sanitizer over the hot path (race detection / sanitizers hook).
#   memcheck  : out-of-bounds / misaligned global and shared accesses, leaks of device memory
#   racecheck : shared-memory hazards (k_linear_memories_band staging, k_coarse_packed task tables + bit-planes,
#               k_refine<true> / k_refine_bits partial sums, block scans)
#   synccheck : barrier / mbarrier misuse (the TMA staging of k_coarse_packed, __syncthreads in divergent code)
# on (a) __graft_entry__.smoke() -- one full match checked against the oracle -- and (b) the fused multi-GPU exchange
# with three handles in one process (peer stores, ticket counters, frame flags, the collector's spin wait), and
# (c) a second match with every alternative path switched on by environment (byte-wise refinement, 8-lane filter).
# Usage (on a GPU box):  bash tools/sanitize.sh [outdir]      -> <outdir>/summary.txt, one log per tool x case
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
export LINEMOD_B200_PEER_TIMEOUT_MS=600000   # kernels run 10-100x slower under the sanitizer: no false peer timeouts
CS=${CS:-compute-sanitizer}
: > "$OUT/summary.txt"
run() {  # tool case command...
  local tool=$1 name=$2; shift 2
  local log="$OUT/${tool}_${name}.log"
  timeout 900 $CS --tool "$tool" --error-exitcode 7 --print-limit 20 --log-file "$log" "$@" > "$OUT/${tool}_${name}.out" 2>&1
  local rc=$?
  local errs
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  echo "$tool $name rc=$rc :: ${errs:-no summary line}" | tee -a "$OUT/summary.txt"
}
for tool in memcheck racecheck synccheck; do
  run "$tool" smoke python -c "import __graft_entry__ as g; g.smoke()"
  run "$tool" exchange3 python -m pytest "tests/test_gpu_dist.py::test_fused_exchange_three_handles_one_process" -x -q -p no:cacheprovider
done
LINEMOD_B200_BITS_EXACT=0 LINEMOD_B200_FILTER_VARIANT=1 LINEMOD_B200_PLANES_DIRECT=0 run memcheck smoke_altpaths python -c "import __graft_entry__ as g; g.smoke()"
LINEMOD_B200_FILTER=0 run memcheck smoke_nofilter python -c "import __graft_entry__ as g; g.smoke()"
echo "done: $(cat "$OUT/summary.txt" | wc -l) runs" | tee -a "$OUT/summary.txt"
