"""Random-draw providers for code that calls the reference-signature functions WITHOUT the extra ``raw_noise=`` / ``u=``
arguments (i.e. the reference's own Trainer._train_step after install()).  By default the kernels draw in-kernel (counter
hashes); a parity test that must feed the same numbers to the reference and to this package sets

    _hooks.normal  = fn(n, device)     -> (n,) fp32 N(0,1) draws      replaces torch.randn at utils/render_utils.py:240
    _hooks.uniform = fn(B, n, device)  -> (B, n) fp32 U[0,1) draws    replaces torch.rand  at utils/sampling_utils.py:48
    _hooks.jitter  = fn(B, nc, device) -> (B, nc) fp32 U[0,1) draws   replaces torch.rand_like at train/trainer.py:907
                                          (consulted by the FUSED train step only: the reference's own inline sampler
                                          calls torch.rand_like itself, which such a test patches directly)

and resets them to None afterwards."""
normal = None
uniform = None
jitter = None
