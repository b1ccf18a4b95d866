#!/bin/bash
# One gpurun call of round 2 (development tool): usage tools/gpu_call.sh <step>...
mkdir -p gpurun_out
for step in "$@"; do
case $step in
probe:*)
  IFS=: read -r _ lib cases <<< "$step"
  echo "=== $lib $cases" >> gpurun_out/probe.log
  AIRS_PROBE_LIB=$lib timeout 600 python tools/perf_probe.py ${cases//,/ } >> gpurun_out/probe.log 2>&1 ;;
tests)
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log ;;
pytest:*)
  IFS=: read -r _ what <<< "$step"
  timeout 2400 python -m pytest ${what//,/ } -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
  tail -n 25 gpurun_out/pytest.log ;;
bench:*)
  IFS=: read -r _ name bargs <<< "$step"
  timeout 1500 python bench.py ${bargs//,/ } > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "bench rc=$?"
  tail -n 5 gpurun_out/$name.err; head -c 6000 gpurun_out/$name.json ;;
bounds)
  timeout 600 python tools/bounds_check.py > gpurun_out/bounds.log 2>&1; echo "bounds rc=$?" >> gpurun_out/bounds.log; tail -n 8 gpurun_out/bounds.log ;;
one:*)
  IFS=: read -r _ what <<< "$step"
  timeout 300 python -m pytest "${what}" -x -q -m gpu > gpurun_out/one.log 2>&1; echo "one rc=$?" >> gpurun_out/one.log
  tail -n 25 gpurun_out/one.log ;;
smoke)
  timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log ;;
ncu:*)
  IFS=: read -r _ case kern name <<< "$step"
  timeout 300 python tools/ncu_case.py $case > gpurun_out/ncu_plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/$name python tools/ncu_case.py $case > gpurun_out/ncu_$name.log 2>&1
  tail -n 3 gpurun_out/ncu_$name.log ;;
launches:*)
  IFS=: read -r _ case name <<< "$step"
  timeout 300 python tools/ncu_case.py $case > gpurun_out/ncu_plain_$name.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:airs\|concat\|Memset -c 40 --csv --log-file gpurun_out/$name.csv python tools/ncu_case.py $case > gpurun_out/ncu_$name.log 2>&1
  tail -n 3 gpurun_out/ncu_$name.log ;;
launchbench:*)   # launch list of the default bench command: launchbench:<name>:<bench args>
  IFS=: read -r _ name bargs <<< "$step"
  timeout 900 python bench.py ${bargs//,/ } > gpurun_out/ncu_plain_$name.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:airs\|concat -c 400 --csv --log-file gpurun_out/$name.csv python bench.py ${bargs//,/ } > gpurun_out/ncu_$name.log 2>&1
  tail -n 2 gpurun_out/ncu_$name.log | cut -c1-300 ;;
ncubench:*)      # full capture of one kernel of a bench command: ncubench:<kernel regex>:<skip>:<name>:<bench args>
  IFS=: read -r _ kern skip name bargs <<< "$step"
  timeout 900 python bench.py ${bargs//,/ } > gpurun_out/ncu_plain_$name.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/$name python bench.py ${bargs//,/ } > gpurun_out/ncu_$name.log 2>&1
  tail -n 2 gpurun_out/ncu_$name.log | cut -c1-300 ;;
ncucase:*)       # full capture of one kernel of a probe case: ncucase:<case>:<kernel regex>:<skip>:<name>
  IFS=: read -r _ case kern skip name <<< "$step"
  timeout 300 python tools/ncu_case.py $case > gpurun_out/ncu_plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/$name python tools/ncu_case.py $case > gpurun_out/ncu_$name.log 2>&1
  tail -n 3 gpurun_out/ncu_$name.log ;;
*) echo "unknown step $step" ;;
esac
done
tail -n 40 gpurun_out/probe.log 2>/dev/null
tail -n 30 gpurun_out/tests.log 2>/dev/null
tail -n 5 gpurun_out/smoke.log 2>/dev/null
