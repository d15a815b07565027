mkdir -p gpurun_out/tb; cd $GRAFT_REPO_ROOT
timeout 500 python profiles/tune.py --config c5 --refetch 1 --iters 25 --combos 16:0:0:0,32:12:2:0,32:14:4:0,32:12:4:0,32:14:2:0,64:12:1:0,64:12:2:0,64:8:2:0 > gpurun_out/tb/tune_c5_i8_blocks.jsonl 2> gpurun_out/tb/err_c5.txt
timeout 600 python profiles/tune.py --config c3 --refetch 1 --iters 20 --combos 64:12:2:0,64:16:2:0,64:10:2:0,64:14:2:0,32:14:4:0 > gpurun_out/tb/tune_c3_i8_blocks.jsonl 2> gpurun_out/tb/err_c3.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/tb/*.jsonl")):
    print(f)
    for l in open(f):
        try:
            d=json.loads(l); g=d.get("geom",{}); print(' ',d["combo"], round(d.get("ms_last10_mean",0),3), g.get("block"), g.get("lookahead"), g.get("tile_stages"), g.get("smem_bytes"), d.get("error","")[:100])
        except Exception as ex: print('  ?', l[:200])
PY
tail -3 gpurun_out/tb/err_c5.txt gpurun_out/tb/err_c3.txt
