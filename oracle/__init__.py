"""oracle/ — CPU restatement of the RiboDetector hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``ribodetector_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and only as the *checker* or as the CPU arm being timed.

What is restated (reference paths are relative to ``/root/reference``):

* ``encoders``      – ``ribodetector/data_loader/seq_encoder.py:11-18,126-145``
* ``model_torch``   – ``ribodetector/model/model.py:10-37,114-119`` (packed semantics) and
                      ``ribodetector/model/model_cpu.py:8-37,57-62`` (padded semantics).  The
                      arithmetic itself lives in a third-party dependency of the reference,
                      PyTorch ``nn.LSTM`` / ``nn.Linear`` (``setup.py:15`` pins
                      ``torch>=1.7.1,<=1.12.1``; this image has 2.11.0+cu128), which is present
                      here, so the restatement calls the same library on CPU.
* ``model_numpy``   – an independent fp64/fp32 NumPy restatement of the published LSTM cell
                      equations incl. the reverse-direction LUT collapse; the
                      arbiter when torch-CPU and the CUDA kernels differ in the last bits.
* ``pairs``         – ``ribodetector/detect.py:601-663`` (label routing / pair combination).
* ``cpu_pipeline``  – the ``ribodetector_cpu`` batch loop ``detect_cpu.py:686-708`` as a
                      torch-CPU stand-in (onnxruntime is not installed in this image).

Pinning: the reference ships NO tests / golden vectors.  The oracle is pinned
against outputs of the reference itself, imported in the build container by
``oracle/gen_golden.py`` (committed) which wrote ``tests/golden/*.npz`` (committed) and
asserts reference == oracle while doing so.
"""
