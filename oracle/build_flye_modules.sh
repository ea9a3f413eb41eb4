#!/bin/bash
# TEST INFRASTRUCTURE.  Builds, from the reference sources where they lie (nothing is copied; build container only):
#   oracle/_ref/flye-modules-ref    the UNMODIFIED reference `flye-modules` (all stages)
#   oracle/_ref/flye-modules-b200   the SAME callers (assemble / repeat_graph / contigger / polishing, alignment.cpp, edlib.cpp,
#                                   consensus_generator.cpp, main.cpp) compiled against the flye_b200 host mirror: a shadow tree of
#                                   symlinks in which src/sequence/{sequence,sequence_container,kmer,vertex_index,overlap}.h are the
#                                   mirror's headers and the four .cpp files they replace are left out; linked to libflye_b200.so.
# This is INTEGRATION.md §1 executed: the drop-in proof is that the callers compile, link and (tests/test_gpu_flye_modules.py)
# produce the reference's draft assembly on the device path.
set -e
R=${REF:-/root/reference}
REPO=$(cd "$(dirname "$0")/.." && pwd)
OUT=$REPO/oracle/_ref; W=$REPO/build/fm
INC="-I$R/lib/libcuckoo -I$R/lib/interval_tree -I$R/lib/lemon -I$R/lib/minimap2"
CXXF="-O3 -DNDEBUG -pthread -fsigned-char -w -include cstdint"
mkdir -p $OUT $W/ref $W/b200
for c in kalloc ksw2_extz2_sse; do [ -f $W/mm_$c.o ] || gcc -c -O2 -w -msse2 -DHAVE_KALLOC -I$R/lib/minimap2 $R/lib/minimap2/$c.c -o $W/mm_$c.o; done
# (1) the reference as it is (flags of src/Makefile:3,12 + `-include cstdint` for g++ 13)
if [ ! -x $OUT/flye-modules-ref ]; then
    (cd $R/src; ls main.cpp */*.cpp) | while read f; do o=$W/ref/$(echo $f | tr / _).o; [ -f $o ] || echo "g++ -c -std=c++11 $CXXF $INC $R/src/$f -o $o"; done > $W/ref_cmds.txt
    xargs -P 8 -I{} bash -c "{}" < $W/ref_cmds.txt
    g++ -pthread $W/ref/*.o $W/mm_*.o -lz -rdynamic -o $OUT/flye-modules-ref
fi
# (2) the callers on the mirror
S=$W/shadow/src; rm -rf $W/shadow; mkdir -p $S
for d in assemble repeat_graph common contigger polishing sequence; do mkdir $S/$d; for f in $R/src/$d/*; do ln -s $f $S/$d/; done; done
ln -s $R/src/main.cpp $S/main.cpp
for f in $REPO/flye_b200/host/sequence/*.h; do ln -sf $f $S/sequence/; done
for f in vertex_index overlap sequence_container sequence; do rm $S/sequence/$f.cpp; done
(cd $S; ls main.cpp */*.cpp) | while read f; do echo "g++ -c -std=c++17 $CXXF $INC -I$REPO/include $S/$f -o $W/b200/$(echo $f | tr / _).o"; done > $W/b200_cmds.txt
xargs -P 8 -I{} bash -c "{}" < $W/b200_cmds.txt
g++ -pthread $W/b200/*.o $W/mm_*.o -L$REPO/flye_b200 -lflye_b200 -lz -rdynamic -Wl,-rpath,'$ORIGIN/../../flye_b200' -o $OUT/flye-modules-b200
echo "built $OUT/flye-modules-ref $OUT/flye-modules-b200"
