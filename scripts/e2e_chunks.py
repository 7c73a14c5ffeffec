"""Development: the facade's host path (compress_batch_channel_latents on pinned NumPy arrays) for several chunk sizes."""
import os, subprocess, sys
for rows in (2304, 4608, 9216, 18432, 36864):
    env = dict(os.environ, VBQ_HOST_CHUNK_ROWS=str(rows))
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "prof_facade.py")], env=env, capture_output=True, text=True).stdout
    print(rows, out.strip().splitlines()[-1])
