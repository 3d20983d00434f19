"""CPU oracle for the EventPretrain input hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product package (eventpretrain_b200/) never does.

* ``oracle.events``   — ctypes binding of the C restatement (oracle/ep_oracle.c) of stage 1
  (voxel grid, count frames, hot-pixel filter, normalisers, EvRep).
* ``oracle.stage3_np`` — numpy restatement of stage 2/3 (diff-map target, masking, patchify, gathers).

Parity status: pinned against golden vectors produced by executing the unmodified reference
(tests/golden/make_golden.py); see tests/test_oracle_golden.py.
"""
