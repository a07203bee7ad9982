#!/bin/bash
# One command that turns "parity unpinned" into "pinned" on a host that has a JVM and a build of the reference:
#
#     tools/jaad_verify/run.sh <classpath of pucgenie/JAADec: its jar, or the build/classes directories joined with ':'>
#
# 1. export_streams.py writes every stream of tests/golden/*.npz as .aac / .mp4 files;
# 2. JaadDump.java (compiled against the given classpath) decodes them with the reference's own front-ends
#    (ADTSDemultiplexer / MP4Container, as net.sourceforge.jaad.Main does) and dumps the float channels of every frame;
# 3. compare_jaad_dump.py requires int16 PCM and float bits identical to the fixtures the oracle produced -- the same
#    fixtures the CUDA engine reproduces bit for bit in tests/test_parity_lc_gpu.py::test_engine_reproduces_committed_golden.
set -e
CP="$1"
[ -n "$CP" ] || { echo "usage: $0 <JAAD classpath>"; exit 2; }
HERE="$(cd "$(dirname "$0")" && pwd)"
WORK="${2:-$HERE/_work}"
python3 "$HERE/export_streams.py" "$WORK"
javac -cp "$CP" -d "$WORK/classes" "$HERE/JaadDump.java"
for f in "$WORK"/*.aac "$WORK"/*.mp4; do
  [ -e "$f" ] || continue
  java -cp "$CP:$WORK/classes" JaadDump "$f" "${f%.*}.dump"
done
python3 "$HERE/compare_jaad_dump.py" "$WORK"
