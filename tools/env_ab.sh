# usage: env_ab.sh <config> <spp> VAR=VALUE ...   — one library, environment variants (fresh process each)
cfg=$1; spp=$2; shift 2
echo "== base"; python tools/ab.py raytracinginoneweekendinrust_b200/lib/libshimmer_b200.so --config $cfg --spp $spp --rounds 2 | tail -1
for kv in "$@"; do echo "== $kv"; env $kv python tools/ab.py raytracinginoneweekendinrust_b200/lib/libshimmer_b200.so --config $cfg --spp $spp --rounds 2 | tail -1; done
