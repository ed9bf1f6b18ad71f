# quick loop: headline parity tests + bench (no cpu baseline / extras) + optional ncu capture (NCU=1)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-q}
(python -m pytest tests/test_gpu_sweep.py tests/test_gpu_extra.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_pytest.log; tail -c 300 gpurun_out/${T}_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("parity"))
PY
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k "regex:k_sweep_(prod2|tab2)" -s 1 -c 1 -f -o gpurun_out/${T}_prod2c python scripts/prof_compact.py > gpurun_out/${T}_ncu_full.log 2>&1; tail -2 gpurun_out/${T}_ncu_full.log
fi
