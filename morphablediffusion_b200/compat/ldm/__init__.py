"""Import-path shim: put `morphablediffusion_b200/compat` ahead of the reference checkout on sys.path and the
reference's own scripts (generate_face.py:11, eval/generate_all_facescape.py:14) import the B200 classes unchanged."""
