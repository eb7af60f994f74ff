cd $GRAFT_REPO_ROOT
python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
THZ_LIB=$GRAFT_REPO_ROOT/variants/libthzdoe_r25b2.so python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
