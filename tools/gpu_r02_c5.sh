cd $GRAFT_REPO_ROOT
python tools/c5_ab.py --tag default
THZ_NO_K2FAST=1 python tools/c5_ab.py --tag general_k2
THZ_T2_LOG2=2 python tools/c5_ab.py --tag t2_blocked4
THZ_T2_LOG2=3 python tools/c5_ab.py --tag t2_blocked8
