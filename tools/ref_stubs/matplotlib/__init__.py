"""Import-only stand-in: the reference imports matplotlib at module top
(system/carrier.py:14, system/perf_monitor.py:14) but only its animation code uses it."""
