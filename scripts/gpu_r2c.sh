set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-r2c}
(python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${T}_pytest.log
(python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_smoke.log; tail -c 600 gpurun_out/${T}_bench.err; head -c 1200 gpurun_out/${T}_bench.json
