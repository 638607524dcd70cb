set -x
export PP_B200_LIB=$PWD/path_planning_pkg_b200/lib/libpp_b200_prof.so
# lone warp: 1 slot, 8 queries of group 0: per-phase cycles of one expansion
python scripts/gpu_probe.py --groups 1 --starts 8 --slots 1 2>&1 | tail -22
# crowded: 16 groups x 64 queries on 1024 slots (7 warps per SM)
python scripts/gpu_probe.py --groups 16 --starts 64 --slots 1024 2>&1 | tail -22
