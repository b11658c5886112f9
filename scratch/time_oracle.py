import time, torch, os
from oracle import sg3
torch.set_num_threads(os.cpu_count())
net = sg3.make_synthesis("T", 1024)
ws = torch.randn(1, 16, 512)
for i in range(2):
    t = time.time(); y = net(ws); print("oracle T 1024 B=1:", time.time() - t, "s", y.shape, flush=True)
