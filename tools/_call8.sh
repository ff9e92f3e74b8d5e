for cfg in "8 1" "16 1" "8 2" "16 2" "8 4" "16 4"; do set -- $cfg; L3C_RGB_CHUNKS=$1 python bench.py --workload rgb_shared --value-only --lanes $2 2>&1 | tail -1 | cut -c60-170; done
