"""Host topology as the GPU box shows it (CPU affinity, NUMA nodes, GPU PCI ids and their NUMA node)."""
import os, glob, torch
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...", "cpu_count", os.cpu_count())
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try: print(n, open(n + "/cpulist").read().strip(), open(n+"/meminfo").read().split("\n")[0])
    except Exception as e: print(n, e)
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try: nn = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
    except Exception as e: nn = repr(e)
    print("gpu", i, bus, "numa", nn)
os.system("nvidia-smi topo -m 2>&1 | head -20")
