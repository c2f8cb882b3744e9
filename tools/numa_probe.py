import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
for d in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(d)
    print(d, p.name, getattr(p, "pci_domain_id", None), getattr(p, "pci_bus_id", None), getattr(p, "pci_device_id", None), bench.gpu_locality(torch, d)[0], len(bench.gpu_locality(torch, d)[1]))
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
os.system("ls /sys/devices/system/node/ | head; nvidia-smi topo -m | head -20")
