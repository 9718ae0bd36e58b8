"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle.cpp).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by nptranscript_b200/."""
