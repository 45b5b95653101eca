import ctypes, os, sys, torch
sys.path.insert(0, "/root/repo")
from autogpc_b200.engine import Engine
eng = Engine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for (M,N,K,batch) in [(8192,8192,8192,1),(6400,6400,512,8)]:
    A = torch.randn(batch,M,K,dtype=torch.float64,device="cuda"); B = torch.randn(batch,N,K,dtype=torch.float64,device="cuda"); C = torch.zeros(batch,M,N,dtype=torch.float64,device="cuda")
    f = lambda: eng.gemm_nt_i8x(A.data_ptr(),B.data_ptr(),C.data_ptr(),M,N,K,1.0,0.0,batch,7,st.cuda_stream)
    f(); torch.cuda.synchronize()
    eng.profile_enable(True); f(); p = eng.profile_read(); eng.profile_enable(False)
    fl = 2.0*M*N*K*batch
    print("dbg=%s M=%d K=%d batch=%d: MMA kernel %.3f ms = %.1f TF-eq" % (os.environ.get("AGPC_OZ_DBG","0"), M, K, batch, p["gemm"]["ms"], fl/p["gemm"]["ms"]/1e9), flush=True)
    del A,B,C
