#!/bin/bash
# Stage the committed HEAD (sources + freshly built library) under ab_old/ (git-ignored, travels with gpurun) so that
# tools/ab_old_new.sh can alternate the committed and the working-tree kernels on ONE GPU box (boxes differ by +-5 %).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
rm -rf /tmp/old_wt "$ROOT/ab_old"
git -C "$ROOT" worktree prune
git -C "$ROOT" worktree add -f /tmp/old_wt HEAD -q
(cd /tmp/old_wt/pcss-unet_b200 && python -c "import build; build.build(verbose=False)")
mkdir -p "$ROOT/ab_old"
(cd /tmp/old_wt && tar cf - --exclude=.git --exclude='*.o' --exclude=profiles --exclude=tests --exclude=gpurun_out .) | (cd "$ROOT/ab_old" && tar xf -)
cp "$ROOT/MEASURED_PEAKS.json" "$ROOT/ab_old/" 2>/dev/null || true
git -C "$ROOT" worktree remove --force /tmp/old_wt
