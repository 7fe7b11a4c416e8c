import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_bench import bench
ALL = (0, 1, 2, 3, 4, 5)
bench(12608, 768, 768, b_mn=True, cfgs=ALL)                                   # out-proj dgrad
bench(12608, 768, 768, bias=True, res=True, f32=True, cfgs=ALL)               # out-proj fwd
bench(768, 768, 12608, a_mn=True, b_mn=True, f32=True, acc=True, cfgs=ALL)    # out-proj wgrad
bench(12608, 2304, 768, bias=True, cfgs=ALL)                                  # QKV fwd
bench(12608, 768, 2304, b_mn=True, cfgs=ALL)                                  # QKV dgrad
bench(12608, 3072, 768, bias=True, act=1, cfgs=ALL)                           # fc1 fwd (no out2)
bench(12608, 768, 3072, bias=True, res=True, f32=True, cfgs=ALL)              # fc2 fwd
bench(3072, 768, 12608, a_mn=True, b_mn=True, f32=True, acc=True, cfgs=ALL)   # fc2 wgrad
bench(768, 3072, 12608, a_mn=True, b_mn=True, f32=True, acc=True, cfgs=ALL)   # fc1 wgrad
bench(2304, 768, 12608, a_mn=True, b_mn=True, f32=True, acc=True, cfgs=ALL)   # QKV wgrad
bench(12608, 3072, 768, b_mn=True, cfgs=ALL)                                  # fc2 dgrad (no aux)
bench(12544, 256, 6912, bias=True, act=2, cfgs=ALL)                           # head conv fwd
bench(12544, 6912, 256, b_mn=True, cfgs=ALL)                                  # head dgrad
bench(256, 6912, 12544, a_mn=True, b_mn=True, f32=True, acc=True, cfgs=ALL)   # head wgrad
bench(18464, 3072, 1024, bias=True, act=1, cfgs=ALL)                          # ViT-L fc1 fwd (B=32, N=577)
bench(18464, 1024, 1024, bias=True, res=True, f32=True, cfgs=ALL)             # ViT-L out-proj fwd
bench(18464, 3072, 1024, bias=True, cfgs=ALL)                                 # ViT-L QKV fwd
bench(32800, 2304, 768, bias=True, cfgs=ALL)                                  # ViT-B @512 B=32 QKV fwd
