"""bruno_b200: B200 (sm_100a) implementation of BRUNO's hot path -- the Real NVP coupling stack and the recurrent
Student-t / Gaussian process -- behind the reference's layer and config surface.  The compute lives in
bruno_b200/lib/libbruno_sm100.so (built from csrc/ by bruno_b200.build); importing the package loads nothing yet."""
