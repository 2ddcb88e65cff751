echo "== PLOC"; python tools/prof_step.py config3 3 --stats 2>&1 | tail -2
echo "== LBVH"; ENVGS_BVH_BUILDER=lbvh python tools/prof_step.py config3 3 --stats 2>&1 | tail -2
