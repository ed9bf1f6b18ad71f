# full round check: GPU tests, smoke, bench (with cpu baseline and extras), ncu launch list of bench.py, ncu full capture of the headline kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-full}
(python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${T}_pytest.log
(python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_sweep_(prod2|tab2)" -s 1 -c 1 -f -o gpurun_out/${T}_headline python scripts/prof_compact.py > gpurun_out/${T}_ncu_full.log 2>&1
mkdir -p /tmp/n
COMPARE=0 ncu --set full --clock-control none --import-source on -k regex:k_sweep_rowc -s 1 -c 1 -f -o /tmp/n/${T}_rowc python scripts/probe_rowc.py > gpurun_out/${T}_ncu_rowc.log 2>&1
python scripts/ncu_summary.py /tmp/n/${T}_rowc.ncu-rep 524812288 gpurun_out/${T}_rowc_ncu_summary.txt > /dev/null
python scripts/ncu_lines.py /tmp/n/${T}_rowc.ncu-rep 40 > gpurun_out/${T}_rowc_ncu_lines.txt
tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_smoke.log; tail -c 600 gpurun_out/${T}_bench.err; head -c 1500 gpurun_out/${T}_bench.json
