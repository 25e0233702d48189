for lib in ab_head.so ab_scan.so; do echo "== $lib"; PWGS_LIB=$PWD/$lib python scripts/mh_big_tree.py 128 1 | tail -1; PWGS_LIB=$PWD/$lib python scripts/mh_big_tree.py 64 1 | tail -2; done
