"""
ORACLE -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference legs).  The product package amt_tools_b200 never imports anything from here.
See oracle/librosa_stages.py for the parity status (CQT/VQT/soxr stages: parity unpinned).
"""
