for spec in "$@"; do echo "== $spec"; MMH_LIB=$PWD/metmhn_b200/$spec python scripts/prof_classes.py 20 10000 | tail -1; done
