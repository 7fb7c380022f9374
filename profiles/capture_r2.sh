#!/bin/bash
# Round-2 profile captures (run on the B200 box under gpurun from the repo root):
#   bash profiles/capture_r2.sh A      (then B, then C: gpurun brings back at most 64 MiB per call, a report is ~8 MiB)
# ncu --set full of every hot kernel on the BASELINE configs[1] block, ONE launch covering the whole block (--wave 32),
# at the headline entropy (kv8d, 0.58 bits/symbol) and at the sweep's top (uniform_signed, 4.1 bits/symbol); plus the
# launch list of the default bench command.  Reports land in gpurun_out/; summaries are made afterwards, without a GPU, by
# profiles/summarize.py, profiles/by_line.py and profiles/traffic.py.
set -u
B="python bench.py --wave 32 --steps 1 --warmup 3 --no-e2e --no-cpu --no-sweep"
N="ncu --set full --clock-control none --import-source on -c 1"
o=gpurun_out
part=${1:-A}
if [ "$part" = A ]; then
$N -k regex:absmax_kernel  -s 3 -o $o/r2_absmax_kv8d         $B --data kv8d > /dev/null 2> $o/r2_cap1.err
B200KV_ENCODE_PATH=legacy $N -k regex:encode_kernel  -s 3 -o $o/r2_encode_kv8d         $B --data kv8d > /dev/null 2> $o/r2_cap2.err
$N -k regex:compact_kernel -s 3 -o $o/r2_compact_kv8d        $B --data kv8d > /dev/null 2> $o/r2_cap3.err
B200KV_DECODE_TABLE=rows $N -k regex:decode_kernel  -s 3 -o $o/r2_decode_kv8d         $B --data kv8d > /dev/null 2> $o/r2_cap4.err
fi
if [ "$part" = B ]; then
B200KV_ENCODE_PATH=tma $N -k regex:encode_tma     -s 3 -o $o/r2_encode_tma_uniform  $B --data uniform_signed > /dev/null 2> $o/r2_cap5.err
B200KV_ENCODE_PATH=legacy $N -k regex:encode_kernel  -s 3 -o $o/r2_encode_uniform  $B --data uniform_signed > /dev/null 2> $o/r2_cap6.err
B200KV_DECODE_TABLE=transposed $N -k regex:decode_kernel -s 3 -o $o/r2_decode_tr_uniform $B --data uniform_signed > /dev/null 2> $o/r2_cap7.err
B200KV_ENCODE_PATH=tma $N -k regex:encode_tma     -s 3 -o $o/r2_encode_tma_kv8d     $B --data kv8d > /dev/null 2> $o/r2_cap8.err
fi
if [ "$part" = C ]; then
# the hash chain on its own (8192 tokens, one chain)
cat > /tmp/hash_once.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from lmcache_b200.cache_engine import sha256_prefix_chain
t = torch.randint(0, 32000, (8192,), device="cuda")
for _ in range(4):
    sha256_prefix_chain(t, 256)
PY
$N -k regex:sha256_chain -s 2 -o $o/r2_sha256_chain python /tmp/hash_once.py > /dev/null 2> $o/r2_cap9.err
# every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu > $o/r2_launches_bench.json 2> $o/r2_cap10.err
fi
ls -la $o/r2_*.ncu-rep | wc -l; du -sh $o
