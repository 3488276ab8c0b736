# A/B: serial class timings + timed evaluations for several builds of the library (MMH_LIB); extra env via VAR=... before the lib name
for spec in "$@"; do echo "== $spec"; env ${spec%%:*} MMH_LIB=$PWD/metmhn_b200/${spec##*:} python scripts/prof_classes.py 25 100000 | tail -1; done
