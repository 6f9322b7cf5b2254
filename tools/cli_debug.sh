#!/bin/bash
# development aid: where does the wall time of our CLIs go?
set -e
D=$(mktemp -d); mkdir -p $D/bin; cp oracle/7z_shim.sh $D/bin/7z; chmod +x $D/bin/7z; export PATH=$D/bin:$PATH
python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, sccg_b200
from sccg_genome_compression_b200 import synth
ref,tgt=synth.local_pair(60_000_000, synth.seed_for(2,5))
def image(seq,h):
    full=seq.size//50*50; b=np.empty((full//50,51),dtype=np.uint8); b[:,:50]=seq[:full].reshape(-1,50); b[:,50]=10
    t=seq[full:].tobytes(); return h+b"\n"+b.tobytes()+(t+b"\n" if t else b"")
open("$D/ref.fa","wb").write(image(ref,b">chrR")); open("$D/tgt.fa","wb").write(image(tgt,b">chrT"))
PY
B=sccg-genome-compression_b200/bin
for i in 1 2; do
  /usr/bin/time -v true 2>/dev/null || true
  s=$(date +%s.%N); $B/compress $D/ref.fa $D/tgt.fa $D/o | tail -3; e=$(date +%s.%N); echo "compress wall $(echo "$e - $s" | bc)"
  s=$(date +%s.%N); $B/decompress $D/o/compressed_genome.txt.7z $D/ref.fa $D/dec | tail -3; e=$(date +%s.%N); echo "decompress wall $(echo "$e - $s" | bc)"
done
s=$(date +%s.%N); python -c "
import ctypes; l=ctypes.CDLL('sccg-genome-compression_b200/libsccg_b200.so'); l.sccg_create.restype=ctypes.c_void_p; import time; t=time.time(); c=l.sccg_create(0); print('sccg_create', round(time.time()-t,3))"; e=$(date +%s.%N); echo "python create wall $(echo "$e - $s" | bc)"
rm -rf $D
