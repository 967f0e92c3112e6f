import sys, os
sys.path.insert(0, "/root/repo")
import torch
from tests.test_gpu_fused_linear import run_case
cases = [("cached 515x3072x768", dict(m=515,k=3072,n=768,a_bit=6,w_bit=6,lsq=False,seed=11,use_code_cache=True)),
         ("nocache 515x3072x768", dict(m=515,k=3072,n=768,a_bit=6,w_bit=6,lsq=False,seed=11,use_code_cache=False)),
         ("cached multi 2048x512", dict(m=148*128+300,k=2048,n=512,a_bit=8,w_bit=8,lsq=False,seed=12,use_code_cache=True)),
         ("cached 100x4096x1024", dict(m=100,k=4096,n=1024,a_bit=6,w_bit=6,lsq=True,seed=13,use_code_cache=True)),
         ("multi persistent", dict(m=148*128*2+77,k=256,n=64,a_bit=6,w_bit=6,lsq=False,seed=5))]
which = int(sys.argv[1])
name, kw = cases[which]
try:
    run_case(**kw); torch.cuda.synchronize(); print("OK  ", name)
except BaseException as e:
    print("FAIL", name, type(e).__name__, str(e)[:300].replace("\n"," | "))
