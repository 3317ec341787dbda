#!/bin/bash
# ncu evidence for one round: launch list of a whole bench run + one `--set full` capture per BASELINE config, each of
# the first timed launch of that config (bench.py --profile-region, cudaProfilerStart/Stop).  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/profile_bench.sh r02'
# Reports land in gpurun_out/; tools/ncu_counts.py turns them into profiles/<round>_*.txt and profiles/ncu_counts.json.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess"
$B > $O/${R}_prof_plain.json 2> $O/${R}_prof_plain.err || { echo "plain bench failed"; tail -5 $O/${R}_prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches.csv $B > $O/${R}_launches.log 2>&1
echo "launch list rc=$?"
for c in ${CONFIGS:-c2 c3 c4 c5}; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/${R}_$c \
      $B --profile-region $c > $O/${R}_ncu_$c.log 2>&1
  echo "ncu $c rc=$?"
done
ls -la $O/*.ncu-rep
