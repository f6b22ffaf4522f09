"""The reference-side binding (lepton_b200/adapter/b200_component_coders.hh: BaseEncoder / BaseDecoder adapters over the
C ABI) must compile against the reference's own headers with the reference's own flags.  Needs /root/reference, so it
runs in the build container only; nothing is executed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lepton")), reason="reference tree only exists in the build container")
def test_adapter_compiles_against_reference_headers(tmp_path):
    tu = tmp_path / "tu.cc"
    tu.write_text('#include "b200_component_coders.hh"\n'
                  "BaseEncoder* make_b200_encoder() { return new B200ComponentEncoder; }\n"
                  "BaseDecoder* make_b200_decoder() { return new B200ComponentDecoder; }\n")
    (tmp_path / "version.hh").write_text('#define GIT_REVISION ""\n')
    inc = [str(tmp_path)] + [os.path.join(REF, p) for p in (
        "src/lepton", "src/vp8/util", "src/vp8/model", "src/vp8/encoder", "src/vp8/decoder",
        "dependencies/brotli/c/include")] + [os.path.join(ROOT, "include"), os.path.join(ROOT, "lepton_b200", "adapter")]
    # flags of the reference build: CMakeLists.txt:76-79 (-std=c++11 -fno-exceptions -fno-rtti), :56, :247, :293
    cmd = ["g++", "-std=c++11", "-fno-exceptions", "-fno-rtti", "-mssse3", "-msse4.2", "-DNDEBUG",
           "-DDEFAULT_ALLOW_PROGRESSIVE", "-DHIGH_MEMORY", "-w", "-c", str(tu), "-o", str(tmp_path / "tu.o")]
    for i in inc:
        cmd += ["-I", i]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # both adapters are concrete (every pure virtual of base_coders.hh:26-65 is implemented) and only C-ABI symbols
    # of the library are referenced
    syms = subprocess.run(["nm", "-u", str(tmp_path / "tu.o")], capture_output=True, text=True).stdout
    used = sorted({l.split()[-1] for l in syms.splitlines() if "lepb200_" in l})
    assert used == ["lepb200_create", "lepb200_decode_images", "lepb200_destroy", "lepb200_encode_images"], used


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lepton")), reason="reference tree only exists in the build container")
def test_reference_cli_links_with_the_adapters_and_has_no_cpu_path():
    """build() links the reference CLI with the adapters in its two factory lines (lepton_b200/adapter/Makefile.plug ->
    oracle/_ref/lepton-b200plug).  The binary must reference the adapters and the C ABI, and without a CUDA device
    both directions must end in the adapter's OS_ERROR exit (33) -- there is no CPU coder behind the boundary any more."""
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "lepton-b200plug")
    assert os.path.exists(exe), "build() did not produce the plug binary"
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    assert "B200ComponentEncoder::encode_chunk" in syms and "B200ComponentDecoder::decode_chunk" in syms
    for f in ("lepb200_create", "lepb200_encode_images", "lepb200_decode_images"):
        assert " U " + f in syms
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    golden = os.path.join(ROOT, "tests", "golden")
    r = subprocess.run([exe, "-unjailed", "-skipverify", os.path.join(golden, "androidcrop.jpg"), "/dev/null"], capture_output=True)
    assert r.returncode == 33, (r.returncode, r.stderr[-500:])
    r = subprocess.run([exe, "-unjailed", "-forceprogressive", os.path.join(golden, "androidcrop.lep"), "/dev/null"], capture_output=True)
    assert r.returncode == 33, (r.returncode, r.stderr[-500:])
