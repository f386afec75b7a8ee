python -m msd_b200.sweep --tiles 27000 --batch 250 2>/dev/null | tail -1
python -m msd_b200.sweep --tiles 27000 2>/dev/null | tail -1
