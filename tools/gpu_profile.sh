#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list of the same command, one --set full capture of the
# two dominant kernels.  Outputs under gpurun_out/.
set -x
R=${1:-r01}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/${R}_host.txt; nproc >> gpurun_out/${R}_host.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1
python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
tail -c 3000 gpurun_out/${R}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${R}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -c 1 -o gpurun_out/${R}_prof_msm_acc -f \
    python bench.py --steps 1 --warmup 1 --log-n-msm 22 --log-n-ntt 16 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${R}_ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 3 -o gpurun_out/${R}_prof_ntt -f \
    python bench.py --steps 1 --warmup 1 --log-n-msm 16 --log-n-ntt 24 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${R}_ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_pair_add -c 2 -o gpurun_out/${R}_prof_pair -f \
    python bench.py --steps 1 --warmup 1 --log-n-msm 22 --log-n-ntt 16 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${R}_ncu_pair.log 2>&1
# keep the copy-back under 64 MiB: export the raw/detail pages on the box and drop the .ncu-rep files
for k in msm_acc ntt pair; do
  ncu -i gpurun_out/${R}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/${R}_prof_$k.ncu-rep --page details 2>/dev/null | grep -vE "^\s*$" | head -400 > gpurun_out/${R}_ncu_$k.details.txt
  rm -f gpurun_out/${R}_prof_$k.ncu-rep
done
ls -la gpurun_out
