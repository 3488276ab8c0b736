# A/B: serial class timings + timed evaluations for several builds of the library (MMH_LIB)
for lib in "$@"; do echo "== $lib"; MMH_LIB=$PWD/metmhn_b200/$lib python scripts/prof_classes.py 25 100000 | tail -1; done
