# cautious: every command under a short timeout (new synchronisation code)
mkdir -p gpurun_out/tsp2; cd $GRAFT_REPO_ROOT
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -k "many_rows" 2>&1 | tail -3 || exit 1
timeout 60 python profiles/tune.py --config c5 --refetch 1 --iters 15 --combos 64:12:1:0,32:12:2:0 > gpurun_out/tsp2/tune_c5_i8_split.jsonl 2> gpurun_out/tsp2/err_c5.txt || { echo C5 failed; exit 1; }
timeout 60 python profiles/tune.py --config c5 --storage 2bit --iters 15 --combos 0:0:0:0,64:12:2:0 --refetch 1 > gpurun_out/tsp2/tune_c5_2bit_split.jsonl 2> gpurun_out/tsp2/err_c5b.txt
timeout 90 python profiles/tune.py --config c3 --refetch 1 --iters 12 --combos 64:12:2:0,64:12:1:0 > gpurun_out/tsp2/tune_c3_i8_split.jsonl 2> gpurun_out/tsp2/err_c3.txt
timeout 90 python profiles/tune.py --config c3 --storage 2bit --refetch 1 --iters 12 --combos 64:12:4:0,64:12:2:0 > gpurun_out/tsp2/tune_c3_2bit_split.jsonl 2> gpurun_out/tsp2/err_c3c.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/tsp2/*.jsonl")):
    print(f)
    for l in open(f):
        try:
            d=json.loads(l); g=d.get("geom",{}); print(' ',d["combo"], round(d.get("ms_last10_mean",0),3), g.get("block"), g.get("lookahead"), g.get("tile_stages"), g.get("smem_bytes"), d.get("error","")[:100])
        except Exception as ex: print('  ?', l[:200])
PY
