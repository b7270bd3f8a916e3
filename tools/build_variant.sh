#!/bin/bash
# tools/build_variant.sh NAME FLAGS... -- an experiment build of the product library next to the default one:
# spear_b200/libspear_particles.NAME.so, loaded with SPR_LIB_VARIANT=NAME (spear_b200/particles.py). The default library is
# rebuilt afterwards so that the tree is left as it was.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SPR_EXTRA_NVCC_FLAGS="$*" python spear_b200/build.py --force > /dev/null
cp spear_b200/libspear_particles.so spear_b200/libspear_particles.$NAME.so
python spear_b200/build.py --force > /dev/null
echo "built spear_b200/libspear_particles.$NAME.so with: $*"
