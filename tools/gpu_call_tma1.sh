set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
O=gpurun_out/r2t_pk_probe.jsonl; : > $O
for cfg in "1024 3" "1024 1" "512 3" "256 3" "500 3" "1000 1" "128 3" "64 3"; do
  for tma in 0 1; do
    VP_X_TMA=$tma timeout 300 python tools/pk_only_probe.py $cfg >> $O 2>> gpurun_out/r2t_pk_probe.err || echo "FAILED $cfg tma=$tma rc=$?"
  done
done
VP_X_L2PROMO=0 timeout 300 python tools/pk_only_probe.py 1024 3 >> $O 2>> gpurun_out/r2t_pk_probe.err
VP_X_L2PROMO=2 timeout 300 python tools/pk_only_probe.py 1024 3 >> $O 2>> gpurun_out/r2t_pk_probe.err
python - <<'P'
import json
for l in open('gpurun_out/r2t_pk_probe.jsonl'):
    d=json.loads(l); s=d['stages_ms']
    print(d['N'], d['ncomp'], 'tma', d['tma'], 'promo', d['promo'], d['nsample_crc32'], d['psum_crc32'], '%.9e'%d['psum_sum'], {k:s[k] for k in s if 'fft_x' in k or 'bin_tiles' in k})
P
tail -5 gpurun_out/r2t_pk_probe.err
