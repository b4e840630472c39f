#!/bin/bash
# phase times + one ncu capture of the SpGEMM fill kernel (27-point 64^3 squared)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AOCLSPARSE_B200_SPGEMM_TRACE=1 python tools/spgemm_bench.py > gpurun_out/spgemm_trace.txt 2>&1
tail -40 gpurun_out/spgemm_trace.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemm_row_kernel -c 4 -o gpurun_out/spgemm_rows -f \
   python -c "
import sys; sys.path.insert(0,'aocl-sparse_b200'); sys.path.insert(0,'tests')
import capi, gen_np
lib=capi.AoclSparse()
rp,col,val=gen_np.stencil(27,64,64,64)
m=len(rp)-1
st,h=lib.create_csr('d',0,m,m,len(col),rp,col,val)
st,c=lib.spmm(111,h,h); print(st)
" > gpurun_out/ncu_spgemm.log 2>&1
ls -la gpurun_out/spgemm_rows.ncu-rep
