"""Per-tile clock64 timeline of the tcgen05 kernel.  Needs the PROFILING build of the library (the shipped one has no such knob):
    FP8B_BUILD_PROFILE=1 python fp8-mps-metal_b200/build.py
    FP8B_LIB=profiles/tools/bin/libfp8_b200_profile.so FP8B_GEMM_STORE=1 [SHAPE=M,K,N] [FP8B_GEMM_CFG=c] python profiles/tools/dbg_gemm.py"""
import ctypes, os, sys
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT+'/fp8-mps-metal_b200'); sys.path.insert(0,ROOT+'/tests')
import torch
from _util import capi
L=capi(); dev=torch.device('cuda',0)
M,K,N=(int(v) for v in os.environ.get('SHAPE','4096,3072,12288').split(','))
g=torch.Generator(device=dev).manual_seed(0)
hi=int(os.environ.get('DATA_HI','120'))
A=torch.randint(0,hi,(M,K),dtype=torch.uint8,device=dev,generator=g); B=torch.randint(0,hi,(N,K),dtype=torch.uint8,device=dev,generator=g)
if os.environ.get('DATA_NAN_FREE','1')=='1' and hi>127:
    A=torch.where((A&0x7F)==0x7F, torch.full_like(A,0x3C), A); B=torch.where((B&0x7F)==0x7F, torch.full_like(B,0x3C), B)
C=torch.empty(M,N,dtype=torch.bfloat16,device=dev); one=torch.full((1,),0.01,device=dev)
P=lambda t: ctypes.c_void_p(t.data_ptr())
for i in range(3):
    if i==2: os.environ['FP8B_GEMM_DEBUG']=str(16+int(os.environ.get('DBG_EXTRA','0')))
    rc=L.fp8b_scaled_mm(P(A),P(B),P(C),2,M,N,K,N,P(one),1,P(one),1,None,0,None,None,0,2,None); assert rc==0, rc
torch.cuda.synchronize()
