# compute-sanitizer racecheck / synccheck over the synchronisation protocols of the LDPC kernel (tools/sanitize_cases.py),
# with and without the chain form, group mode at frames = 2 x resident CTAs of a 148-SM part with 4 CTAs per SM... bounded
# by SANITIZE_GROUP_FRAMES.  Run on the GPU box: bash tools/run_sanitizers.sh > gpurun_out/sanitizer.log 2>&1
for tool in racecheck synccheck; do
  for env in "" "DVBS2B200_CHAIN=0"; do
    echo "== $tool $env"
    (time env $env timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py pair level short c34 group) 2>&1 | tail -12
  done
done
echo "== memcheck"
(time timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py pair level short c34 group) 2>&1 | tail -12
echo "== racecheck group mode, SANITIZE_GROUP_FRAMES=1184 (two rounds of 592 resident CTAs)"
(time env SANITIZE_GROUP_FRAMES=1184 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_cases.py group) 2>&1 | tail -8
for env in "" "DVBS2B200_CHAIN=0"; do
  for tool in racecheck synccheck; do
    echo "== $tool $env on tests/test_gpu_golden.py::test_every_table_four_iterations (all 57 LDPC tables, compared with the oracle)"
    (time env $env timeout 1500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_golden.py::test_every_table_four_iterations -q -m gpu) 2>&1 | tail -9
  done
done
