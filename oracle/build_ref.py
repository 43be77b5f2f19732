"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference hot-path file for the
GPU box.

``/root/reference`` exists only in the build container.  The reference's implementation of
the path is one Python file that imports two upstream symbols (``oracle/ref_loader.py`` supplies
them as a 30-line stub ``mmdet``), so "building" the reference for this path means placing a
byte-identical copy of

    /root/reference/mmdet3d_gaussian/models/losses/gaussian_distance_loss.py

under ``oracle/_ref/`` (git-ignored: it never enters the history; NOT gpurun-ignored: it
travels with the snapshot like the built ``.so`` files) together with a manifest holding its
SHA-256.  ``ref_loader.load_reference()`` then finds it on the GPU box, and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs time THAT file (``kind: "reference"``) instead of
the restatement in ``gd_oracle.py``.  Run by ``__graft_entry__.build()`` whenever
``/root/reference`` is present; a no-op otherwise.

    python oracle/build_ref.py
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')
REL = os.path.join('mmdet3d_gaussian', 'models', 'losses', 'gaussian_distance_loss.py')
STAGED = os.path.join(OUT_DIR, 'gaussian_distance_loss.py')
MANIFEST = os.path.join(OUT_DIR, 'MANIFEST.json')


def sha256(path):
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(reference_root='/root/reference'):
    """Returns the staged path, or None when the reference checkout is not present."""
    src = os.path.join(reference_root, REL)
    if not os.path.isfile(src):
        return STAGED if os.path.isfile(STAGED) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = STAGED + f'.tmp{os.getpid()}'
    shutil.copyfile(src, tmp)
    os.replace(tmp, STAGED)
    with open(MANIFEST + '.tmp', 'w') as f:
        json.dump({'source': src, 'sha256': sha256(STAGED), 'bytes': os.path.getsize(STAGED),
                   'note': 'byte-identical copy of the reference file; not part of the repo'}, f,
                  indent=1)
    os.replace(MANIFEST + '.tmp', MANIFEST)
    return STAGED


def staged_is_intact():
    try:
        with open(MANIFEST) as f:
            return json.load(f)['sha256'] == sha256(STAGED)
    except Exception:
        return False


if __name__ == '__main__':
    path = build(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print(path, 'intact' if path and staged_is_intact() else 'missing')
