#!/bin/bash
# compute-sanitizer passes of round 2 (one B200): memcheck over every kernel, racecheck over the pair-sum kernel's
# mbarrier/atomic stage ring in the fine AND the coarse configuration (O2D_PP_FINE_BELOW=0 forces coarse at small sizes)
# and over the warp-per-particle closest-panel search.
cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck, pp part";   O2D_SANITIZE_PART=pp   $CS --tool memcheck  python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|SANITIZE_SMOKE|Error|Invalid" | head
echo "== memcheck, pp part, coarse configuration forced"; O2D_PP_FINE_BELOW=0 O2D_SANITIZE_PART=pp $CS --tool memcheck python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|SANITIZE_SMOKE|Error|Invalid" | head
echo "== memcheck, rest";      O2D_SANITIZE_PART=rest $CS --tool memcheck  python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|SANITIZE_SMOKE|Error|Invalid" | head
echo "== racecheck, tiny (fine configuration)";   O2D_SANITIZE_PART=tiny $CS --tool racecheck python tools/sanitize_smoke.py 2>&1 | grep -E "RACECHECK SUMMARY|SANITIZE_SMOKE|hazard" | head
echo "== racecheck, tiny, coarse configuration forced"; O2D_PP_FINE_BELOW=0 O2D_SANITIZE_PART=tiny $CS --tool racecheck python tools/sanitize_smoke.py 2>&1 | grep -E "RACECHECK SUMMARY|SANITIZE_SMOKE|hazard" | head
echo "== synccheck, tiny, coarse configuration forced"; O2D_PP_FINE_BELOW=0 O2D_SANITIZE_PART=tiny $CS --tool synccheck python tools/sanitize_smoke.py 2>&1 | grep -E "ERROR SUMMARY|SANITIZE_SMOKE|Error" | head
} | tee gpurun_out/r02_sanitizer.txt
