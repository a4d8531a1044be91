import torch, time, sys
sys.path.insert(0, "/root/repo")
from as_cops_and_thieves_b200 import mappo
dev = "cuda"
N = 4096
pol = mappo.LSTMPolicyNet().to(dev); val = mappo.LSTMValueNet().to(dev)
obs = torch.rand(N, 1, 180, device=dev); st = torch.rand(N, 1, 1090, device=dev)
hc = pol.initial_state(N, dev); hv = val.initial_state(N, dev)
reset = torch.zeros(N, 1, dtype=torch.bool, device=dev)
def t(fn, name, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n*1e3:9.1f} us")
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32; torch.backends.cudnn.allow_tf32 = tf32
    print("tf32", tf32)
    with torch.no_grad():
        t(lambda: pol.features_extractor(obs.reshape(N, 2, 90)), "policy features (conv,conv,linear)")
        t(lambda: pol.features_extractor[0](obs.reshape(N, 2, 90)), "  conv1")
        x1 = pol.features_extractor[1](pol.features_extractor[0](obs.reshape(N, 2, 90)))
        t(lambda: pol.features_extractor[2](x1), "  conv2")
        f = torch.rand(N, 1, 256, device=dev)
        t(lambda: pol.lstm(f, hc), "  lstm 1 step")
        t(lambda: pol(obs, hc, reset), "policy forward")
        t(lambda: val(st, hv, reset), "value forward")
        t(lambda: val.critic_channels(st), "  critic channels")
    # training minibatch
    B, L = 4096, 16
    ob = torch.rand(B, L, 180, device=dev); rs = torch.zeros(B, L, dtype=torch.bool, device=dev)
    h0 = pol.initial_state(B, dev)
    def train():
        lg, _ = pol(ob, h0, rs); lg.sum().backward()
    t(train, "policy fwd+bwd 4096x16", 5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t(train, "policy fwd+bwd 4096x16 bf16 autocast", 5)
