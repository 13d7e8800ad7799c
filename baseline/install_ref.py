"""baseline/install_ref.py -- BUILD CONTAINER ONLY.  Copies the Python sources of the reference that its own render / train
path needs (models/, encoder/ without the CUDA build trees, utils/ray_utils.py, utils/constant.py, geometry/pcd_projector.py) from /root/reference into
baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box), unmodified, so that
`bench.py --impl reference` and the `reference_gpu_path` leg can run the REFERENCE'S OWN NeRFRenderer.run / NeRFNetwork.
Nothing under baseline/_ref is product source or tracked by git.    python baseline/install_ref.py"""
import os
import shutil
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["models/__init__.py", "models/instant_nsr.py", "encoder/__init__.py", "encoder/freq_encoder.py",
         "encoder/hashencoder/__init__.py", "encoder/hashencoder/hashgrid.py", "encoder/shencoder/__init__.py",
         "encoder/shencoder/sphere_harmonics.py", "utils/__init__.py", "utils/ray_utils.py", "utils/constant.py", "geometry/__init__.py", "geometry/pcd_projector.py"]


def main():
    if not os.path.isdir(REF):
        print("install_ref: /root/reference not present, skipping")
        return 0
    n = 0
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst); n += 1
        elif rel.endswith("__init__.py"):
            open(dst, "a").close()                      # namespace marker the reference leaves implicit
    print(f"install_ref: {n} reference files under {OUT}")
    return 0


if __name__ == "__main__":
    main()               # no sys.exit: __graft_entry__.build() runs this file through runpy
