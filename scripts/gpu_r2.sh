#!/bin/bash
# One parameterised runner for the round-2 GPU calls: scripts/gpu_r2.sh <tag> <step> [<step> ...]; every step logs to
# gpurun_out/<tag>_<step>.log with its exit code on the last line. Steps:
#   pytest | smoke | bench | benchref | bolddiag | variants | ncusgd | hosttest | <anything else>: run as a command
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  log=gpurun_out/${tag}_${step//[^a-zA-Z0-9]/_}.log
  case "$step" in
    pytest)   ( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $log 2>&1 ;;
    smoke)    ( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $log 2>&1 ;;
    bench)    ( time timeout 900 python bench.py ) > $log 2>&1 ;;
    benchref) ( time timeout 600 python bench.py --impl reference ) > $log 2>&1 ;;
    bolddiag) ( timeout 300 python scripts/diag_bold_driver.py 0.01 0.6 ) > $log 2>&1 ;;
    variants) ( timeout 600 python scripts/sweep_groups.py --workload netflix --epochs 6 --variants 0,1,2,4 --shapes 37x4 &&
                timeout 300 python scripts/sweep_groups.py --workload netflix --epochs 6 --variants 3 --shapes 37x8,74x4 &&
                timeout 300 python scripts/sweep_groups.py --workload ml10m --epochs 6 --variants 0,1,2,4 --shapes 37x4,9x16 &&
                timeout 300 python scripts/sweep_groups.py --workload ml10m --epochs 6 --variants 3 --shapes 37x8,18x16 ) > $log 2>&1 ;;
    ncusgd)   # launch list + one full capture of the epoch kernel of the default bench workload (after a plain run of the same command)
              ( CMD="python bench.py --sgd-only --no-cpu --steps 2 --warmup 3"
                timeout 300 $CMD > gpurun_out/${tag}_ncusgd_plain.log 2>&1 &&
                timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sgd_|evaluate_|strata_|rs_|scan_|hist|init_rows|stats_' --csv --log-file gpurun_out/${tag}_sgd_launches.csv $CMD &&
                timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_epoch -s 3 -c 1 -f -o gpurun_out/${tag}_sgd_prof $CMD ) > $log 2>&1 ;;
    hostmulti) ( LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib:$LD_LIBRARY_PATH timeout 300 tests/cpp/build/host_test multi 2 10000000 ) > $log 2>&1 ;;
    topntrace) ( MMLB200_TRACE=1 timeout 300 python scripts/bench_topn.py --reps 2 --exact-users 256 ) > $log 2>&1 ;;
    waitstats) ( for shape in 37x4 18x8 9x16 4x37; do MMLB200_SGD_WAITSTATS=1 timeout 200 python scripts/sweep_groups.py --workload netflix --epochs 3 --shapes $shape 2>&1 | cut -c1-160 | tail -n 2; done
                 for shape in 37x4 18x8 9x16 4x37 2x74; do MMLB200_SGD_WAITSTATS=1 timeout 200 python scripts/sweep_groups.py --workload nf_sub8 --epochs 3 --shapes $shape 2>&1 | cut -c1-160 | tail -n 2; done ) > $log 2>&1 ;;
    wrmfdiag) ( timeout 600 python scripts/diag_wrmf_solvers.py --epochs 4 ) > $log 2>&1 ;;
    wrmfncu)  ( CMD="python scripts/bench_wrmf.py --epochs 2"
                timeout 300 $CMD > gpurun_out/${tag}_wrmf_plain.log 2>&1 &&
                timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'wrmf_|gram_' --csv --log-file gpurun_out/${tag}_wrmf_launches.csv $CMD ) > $log 2>&1 ;;
    memcheck)  ( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_case.py ) > $log 2>&1 ;;
    racecheck) ( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_case.py ) > $log 2>&1 ;;
    hosttest) ( LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:$LD_LIBRARY_PATH timeout 60 tests/cpp/build/host_test gpu tests/golden/example.train tests/golden/example.test gpurun_out ) > $log 2>&1 ;;
    *)        ( eval "timeout 900 env $step" ) > $log 2>&1 ;;
  esac
  echo "rc=$?" >> $log
  echo "== $step: $(tail -n 1 $log)"; tail -n 4 $log | head -n 3
done
