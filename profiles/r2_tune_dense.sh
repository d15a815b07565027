mkdir -p gpurun_out/td; cd $GRAFT_REPO_ROOT
timeout 150 python profiles/tune.py --config c4rr --iters 12 --combos 0:0:0:0,64:1:0:1,64:2:0:2,64:3:0:3,64:4:0:4,64:6:0:0,64:6:0:6,32:3:0:3,32:6:0:6,16:4:0:4,16:8:0:8 > gpurun_out/td/tune_c4rr_dense.jsonl 2> gpurun_out/td/err.txt
timeout 60 python profiles/tune.py --config c1 --iters 30 --combos 0:0:0:0,64:1:0:1,64:2:0:2,64:3:0:3,64:4:0:4,32:3:0:3 > gpurun_out/td/tune_c1_dense.jsonl 2>> gpurun_out/td/err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/td/*.jsonl")):
    print(f)
    for l in open(f):
        try:
            d=json.loads(l); g=d.get("geom",{}); print(' ',d["combo"], round(d.get("ms_last10_mean",0),3), g.get("block"), g.get("lookahead"), g.get("near_depth"), g.get("tile_stages"), d.get("error","")[:100])
        except Exception as ex: print('  ?', l[:200])
PY
tail -2 gpurun_out/td/err.txt
