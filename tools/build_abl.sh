#!/bin/bash
# timing-only ablation builds of the nii kernel (tools/scratch/abl/*.so, git-ignored): each drops one ingredient of the "in" pass
cd "$(dirname "$0")/.."
SRC="api.cu decode_quad.cu decode_tpf.cu decode_nii.cu encode.cu modem.cu waveform.cu microbench.cu"
for f in NOCHAN NOGATHER NOY NOCK NOREC NORAW NOPREP ALL; do
  if [ $f = ALL ]; then D="-DNII_ABL_NOCHAN -DNII_ABL_NOGATHER -DNII_ABL_NOY -DNII_ABL_NOCK -DNII_ABL_NOREC -DNII_ABL_NORAW -DNII_ABL_NOPREP"; else D="-DNII_ABL_$f"; fi
  ( cd modulations_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC --fmad=false -ccbin g++ $D -o ../../tools/scratch/abl/lib_$f.so $SRC ) &
done
wait
ls -la tools/scratch/abl/
