"""tools/dram_write_probe.py — how many DRAM bytes does a pure streaming WRITE read?  (ncu: dram__bytes_read / dram__bytes_write of
a 4 GiB fill and of a 4 GiB copy.)  Answers where the config-3 kernel's "extra" DRAM reads come from."""
import torch
x = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
y = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
for _ in range(3):
    x.fill_(7)          # pure write, 4 GiB
    y.copy_(x)          # read 4 GiB + write 4 GiB
torch.cuda.synchronize()
print("ok")
