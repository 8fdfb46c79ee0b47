for il in 0 3 0 3; do echo "IL=$il"; DS2_FLASH_IL=$il python bench.py --only-device --no-cpu-baseline --no-extra-legs 2>/dev/null | cut -c1-120; done
