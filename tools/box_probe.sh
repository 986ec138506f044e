#!/bin/bash
# Facts about the GPU box's host side that the e2e numbers depend on.
echo "kernel: $(uname -r)"; echo "cpus: $(nproc)"; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
echo "thp enabled: $(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)"; echo "thp defrag: $(cat /sys/kernel/mm/transparent_hugepage/defrag 2>/dev/null)"
free -g | head -2; df -h /dev/shm /tmp | cat; nvidia-smi topo -m 2>/dev/null | head -20
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv 2>/dev/null
