for v in 0 1; do
  if [ $v = 1 ]; then export BT_WIDE_CTAS=1; else unset BT_WIDE_CTAS; fi
  echo "BT_WIDE_CTAS=$v"; python tools/prof_target.py 2>&1 | tail -6 | head -3
  python tools/sweep_regen.py --one
done
