"""Prints what the box exposes about GPU / CPU / NUMA placement (used to interpret multi-GPU e2e numbers)."""
import glob
import os
import subprocess

print(subprocess.run("nvidia-smi topo -m | head -12", shell=True, capture_output=True, text=True).stdout)
print("nodes:", sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node*")))
print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
try:
    import pynvml as N
    N.nvmlInit()
    for i in range(N.nvmlDeviceGetCount()):
        h = N.nvmlDeviceGetHandleByIndex(i)
        bus = N.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = open(f"/sys/bus/pci/devices/{bus.lower()[4:] if len(bus) > 12 else bus.lower()}/numa_node").read().strip() \
            if os.path.exists(f"/sys/bus/pci/devices/{bus.lower()[4:] if len(bus) > 12 else bus.lower()}/numa_node") else "?"
        print(i, bus, "numa_node", node, "nvml cpu affinity", [hex(x) for x in N.nvmlDeviceGetCpuAffinity(h, 2)])
except Exception as e:  # noqa: BLE001
    print("nvml:", e)
