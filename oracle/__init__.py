"""ORACLE -- CPU restatement of the reference path.  Test infrastructure only: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from vcf2prot_b200/."""
