"""Minimal stand-in for the 12 MONAI 0.4.0 symbols the reference's network/loss files import.

TEST TOOL ONLY: lets oracle/make_golden.py import /root/reference/params/{networks,losses}
unmodified in the build container (MONAI is not installable offline).  Never imported by the
product code, the GPU tests, smoke() or bench.py.
"""
