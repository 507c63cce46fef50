# Cumulative cost of the phases of sp_relax_exact_kernel (profiling aid CARTB200_SP_PHASE):
#   1 = stop after staging the label tiles, 2 = stop after the border list, 0 = the whole iteration
for ph in 1 2 0; do echo "phase $ph"; CARTB200_SP_PHASE=$ph python tools/stage_bench.py --batch 16 --tag _ph$ph 2>&1 | grep "sp_relax_per"; done
