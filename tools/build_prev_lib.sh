#!/bin/bash
# Build the library of another commit next to the current one (cutmix_semisup_seg_b200/libb200seg_prev.so, git-ignored) for the
# same-box A/B stages of tools/gpu_session.sh (libab, epi2slot): bench.py / the tools load it through B200SEG_LIB.
#   bash tools/build_prev_lib.sh [commit]        (default: HEAD, i.e. the state before the uncommitted change under test)
set -e
rev=${1:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" worktree add --force "$tmp/w" "$rev" > /dev/null
( cd "$tmp/w" && python -c "
import sys; sys.path.insert(0, '.')
from cutmix_semisup_seg_b200 import build
print(build.build_lib())" )
cp "$tmp/w/cutmix_semisup_seg_b200/libb200seg.so" "$root/cutmix_semisup_seg_b200/libb200seg_prev.so"
git -C "$root" worktree remove --force "$tmp/w"
git -C "$root" worktree prune
rm -rf "$tmp"
echo "built $root/cutmix_semisup_seg_b200/libb200seg_prev.so from $rev"
