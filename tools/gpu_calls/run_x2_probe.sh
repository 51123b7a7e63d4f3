cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_pairs or headline_shape or empty_and_tiny" > gpurun_out/r02x2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02x2a_pytest.log
timeout 200 python bench.py --headline-only --skip-cpu --steps 3 --warmup 3 > gpurun_out/r02x2a_bench_on.json 2> gpurun_out/r02x2a_bench_on.err; echo "bench on rc=$?"
POYB200_CONFIG=pair2=0 timeout 200 python bench.py --headline-only --skip-cpu --steps 3 --warmup 3 > gpurun_out/r02x2a_bench_off.json 2> gpurun_out/r02x2a_bench_off.err; echo "bench off rc=$?"
python - <<'PY'
import json
for f in ("on","off"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02x2a_bench_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("phase_ms"), d["e2e"]["value"], d.get("cost_checksum"))
    except Exception as e: print(f, "ERR", e)
PY
