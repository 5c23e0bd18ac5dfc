for l in runs staged; do echo "FLK_LAYOUT=$l"; FLK_LAYOUT=$l python tools/time_configs.py n60000 n250000 1m n4000000 weak16m; done
