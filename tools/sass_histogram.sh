#!/bin/bash
# Opcode histogram of the shipped sm_100a objects (no GPU needed): the Blackwell-specific instructions the design
# relies on, per object.  usage: tools/sass_histogram.sh > profiles/r02_sass_histogram.txt
cd "$(dirname "$0")/../gguf_b200/csrc/build" || exit 1
echo "# cuobjdump -sass of gguf_b200/csrc/build/*.o (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a)"
echo "# UBLKCP = 1-D bulk async copy (TMA engine); SYNCS = mbarrier; LDGSTS = cp.async; FADD2/FMUL2/FFMA2 = packed FP32 (Blackwell);"
echo "# FRND = cvt.rni.f32 on the XU pipe; VHMNMX = 3-input packed half min/max; UTC*MMA / tcgen05: none expected (no contraction on this path)"
printf "%-16s %8s %8s %10s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object kernels UBLKCP.S.G UBLKCP.G.S SYNCS LDGSTS FADD2 FMUL2 FFMA2 FRND VHMNMX HMNMX2 UTCMMA
for o in api.o dequant.o quant_legacy.o quant_k.o rearrange.o; do
  s=$(cuobjdump -sass $o 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-16s %8s %8s %10s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" $o "$(c 'Function :')" "$(c 'UBLKCP.S.G')" "$(c 'UBLKCP.G.S')" "$(c 'SYNCS')" "$(c 'LDGSTS')" "$(c 'FADD2')" "$(c 'FMUL2')" "$(c 'FFMA2')" "$(c ' FRND')" "$(c 'VHMNMX')" "$(c 'HMNMX2')" "$(c 'UTC.*MMA')"
done
echo "# every FFMA2 in quant_k.o multiplies by the kernel argument \`one\` (uniform register): $(cuobjdump -sass quant_k.o | grep FFMA2 | grep -c 'UR[0-9]*\.F32') of $(cuobjdump -sass quant_k.o | grep -c FFMA2)"
echo "# cubins: $(for o in dequant.o quant_legacy.o quant_k.o rearrange.o; do cuobjdump -lelf $o 2>/dev/null | grep -o '[a-z_]*\.sm_[0-9a-z]*\.cubin'; done | tr '\n' ' ')"
