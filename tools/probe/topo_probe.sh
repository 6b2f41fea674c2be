nvidia-smi topo -m 2>&1 | head -14
echo "--- numa"; ls /sys/devices/system/node/ 2>&1 | head; cat /sys/devices/system/node/online 2>&1; cat /sys/devices/system/node/node*/cpulist 2>&1
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo "$d numa=$(cat /sys/bus/pci/devices/$d/numa_node 2>&1)"; done
echo "--- cpus"; nproc; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"
which numactl; python -c "import ctypes; print(ctypes.CDLL('libnuma.so.1'))" 2>&1 | tail -1
free -g | head -2
