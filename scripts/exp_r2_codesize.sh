set -x
python scripts/exp_r2_slots.py --slots 148,592,2368
PP_B200_LIB=$PWD/path_planning_pkg_b200/lib/libpp_b200_spec.so python scripts/exp_r2_slots.py --slots 148,592,2368
