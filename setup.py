"""Packaging surface of the reference (setup.py:41-46): the two console scripts.  The CUDA library
is built in-tree first: python -c "import __graft_entry__ as g; g.build()"."""
from setuptools import setup, find_packages

setup(
    name="ribodetector_b200",
    version="0.1.0",
    packages=find_packages(include=["ribodetector_b200", "ribodetector_b200.*"]),
    package_data={"ribodetector_b200": ["librd_b200.so", "config.json", "data/*.npz"]},
    entry_points={"console_scripts": ["ribodetector=ribodetector_b200.detect:main",
                                      "ribodetector_cpu=ribodetector_b200.detect_cpu:main"]},
)
