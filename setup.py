"""pip install . : compiles libsuchtree_b200.so for sm_100a with nvcc (suchtree_b200/build.py)
and ships it, the CUDA sources and the C header inside the package."""
import os
import shutil

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildWithNvcc(build_py):
    def run(self):
        import importlib.util

        spec = importlib.util.spec_from_file_location("st_build", os.path.join(HERE, "suchtree_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
        # the header travels with the package so that csrc/ can be rebuilt where it is installed
        dst = os.path.join(HERE, "suchtree_b200", "include")
        os.makedirs(dst, exist_ok=True)
        shutil.copyfile(os.path.join(HERE, "include", "suchtree_b200.h"), os.path.join(dst, "suchtree_b200.h"))
        super().run()


setup(
    name="suchtree-b200",
    version="0.1.0",
    description="B200-native batched patristic distances: a drop-in for the hot path of SuchTree",
    packages=["suchtree_b200"],
    package_data={"suchtree_b200": ["libsuchtree_b200.so", "libsuchtree_b200_py.so", "csrc/*.cu", "csrc/*.cuh", "include/*.h",
                                   "pyglue/*.c"]},
    python_requires=">=3.9",
    install_requires=["numpy"],
    cmdclass={"build_py": BuildWithNvcc},
)
