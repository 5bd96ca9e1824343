"""Regenerates tests/golden/ from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box).

  TestIFile_concatenated_compressed.bin  -- binary test resource of the reference
      (tez-runtime-library/src/test/resources/), the only byte-level IFile golden the reference ships.
      It is DATA (5 concatenated zlib-compressed IFile segments written by the real IFile.Writer),
      asserted by RLT/common/sort/impl/TestIFile.java:389-435.
"""
import hashlib
import os
import shutil

SRC = "/root/reference/tez-runtime-library/src/test/resources/TestIFile_concatenated_compressed.bin"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    dst = os.path.join(HERE, os.path.basename(SRC))
    shutil.copyfile(SRC, dst)
    h = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    open(os.path.join(HERE, "SHA256SUMS"), "w").write("%s  %s\n" % (h, os.path.basename(SRC)))
    print(h, os.path.getsize(dst))
