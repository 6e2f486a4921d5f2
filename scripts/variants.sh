for f in metrocpp_b200/variants/*.so; do echo $f; METRO_B200_LIB=$PWD/$f timeout 200 python scripts/partition_bench.py 1000000000 2>&1 | grep "shift 22"; done
