"""Golden outputs of the reference's secondary consumers, produced by RUNNING THE REFERENCE ITSELF:

    python tests/golden/make_golden_consumers.py           (spawns itself in the reference environment)

In the child process `import sige` is the reference's package from baseline/_ref (its python + its sige.cpu kernels); the
models are the reference's own SIGEUNetModel (Stable Diffusion) and SIGEFusedSPADEGenerator (GauGAN) in miniature
(baseline/consumers.py).  Writes tests/golden/sd_mini_golden.npz and gaugan_mini_golden.npz.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def child():
    import numpy as np
    import torch

    import sige

    assert os.path.realpath(sige.__file__).startswith(os.path.realpath(os.path.join(REPO, "baseline", "_ref"))), sige.__file__
    from sige.utils import dilate_mask, downsample_mask

    sys.path.insert(0, os.path.join(REPO, "baseline"))
    import consumers

    torch.set_num_threads(8)
    full0, sparse1 = consumers.run_sd(consumers.build_sd_mini(), downsample_mask)
    np.savez_compressed(os.path.join(HERE, "sd_mini_golden.npz"), full0=full0.numpy(), sparse1=sparse1.numpy())
    print("sd mini:", tuple(sparse1.shape), float(sparse1.abs().max()))
    full0, sparse1 = consumers.run_gaugan(consumers.build_gaugan_mini(), downsample_mask, dilate_mask)
    np.savez_compressed(os.path.join(HERE, "gaugan_mini_golden.npz"), full0=full0.numpy(), sparse1=sparse1.numpy())
    print("gaugan mini:", tuple(sparse1.shape), float(sparse1.abs().max()))


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        sys.path.insert(0, os.path.join(REPO, "baseline"))
        import loader

        sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=loader.reference_env(8)).returncode)
