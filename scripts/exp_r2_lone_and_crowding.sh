set -x
export PP_B200_LIB=$PWD/path_planning_pkg_b200/lib/libpp_b200_prof.so
# lone warp: 1 slot, 8 queries of group 0
python scripts/gpu_probe.py --groups 1 --starts 8 --slots 1 2>&1 | tail -14
# crowding sweep on 512 queries
for s in 148 296 592 1184 2368; do python scripts/gpu_probe.py --groups 8 --starts 64 --slots $s 2>&1 | grep -E "rep 0|cycles/expansion"; done
