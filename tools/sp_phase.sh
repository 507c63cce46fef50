for ph in 3 4 5 0; do echo "phase $ph"; CARTB200_SP_PHASE=$ph python tools/stage_bench.py --batch 16 --tag _ph$ph 2>&1 | grep "sp_relax_per"; done
