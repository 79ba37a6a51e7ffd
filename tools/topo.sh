#!/bin/bash
# host topology of the GPU box (PCIe / NUMA), for reading the multi-GPU end-to-end numbers
nvidia-smi topo -m
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$(basename $d) numa $(cat $d/numa_node) link $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
