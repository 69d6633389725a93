#!/bin/bash
# Build the library of another commit next to the current one (oem_b200/lib/liboem_b200_<tag>.so) for same-box A/B runs:
#   tools/ab_build.sh <commit> <tag>;  OEMB200_LIB_PATH=oem_b200/lib/liboem_b200_<tag>.so python bench.py ...
set -e
commit=$1; tag=$2
root=$(git rev-parse --show-toplevel)
tree=$root/.ab_tree
rm -rf "$tree"; git worktree prune
git worktree add --detach "$tree" "$commit" > /dev/null
( cd "$tree" && python -m oem_b200.build > /dev/null )
cp "$tree/oem_b200/lib/liboem_b200.so" "$root/oem_b200/lib/liboem_b200_$tag.so"
git worktree remove --force "$tree"
echo "$root/oem_b200/lib/liboem_b200_$tag.so"
