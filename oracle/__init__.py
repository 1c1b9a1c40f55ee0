"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's slide-inference path (torch_ref.py), the shim that imports the
reference's own files unmodified where /root/reference exists (ref_shim.py), and the golden-vector
generator (make_golden.py). Nothing under vfmseg_b200/ may import this package.
"""
