"""Generates tests/golden/*.json from the reference tree (run in the build container, where /root/reference exists).

The reference's own tests hold known-answer vectors only for the checksums and two foreign-compressor zip fixtures
(SURVEY.md 8c).  This script extracts them verbatim so that the GPU box (which has no /root/reference) can test
against them:
  * Checksum/ChecksumTests.cs      -> checksum KATs (values copied from the asserts)
  * Zip/ZipCorruptionHandling.cs   -> the raw deflate payloads of TestFileBadCDGoodCD64 and TestFileZeroCodeLength
  * Zip/ZipEncryptionHandling.cs   -> the AES-256 encrypted archive of ZipFileAESReadWithEmptyPassword (:452-482) and the text it
                                      must decrypt + inflate to (pins the oracle's AES-CTR / PBKDF2 / HMAC-SHA1)
"""
import base64
import json
import os
import re
import struct

REF = "/root/reference/test/ICSharpCode.SharpZipLib.Tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def const_string(src, name):
    m = re.search(r"const string %s\s*=\s*((?:\s*@?\"[^\"]*\"\s*\+?)+);" % name, src)
    parts = re.findall(r"\"([^\"]*)\"", m.group(1))
    return "".join(parts)


def first_entry_payload(zipbytes):
    sig, ver, flags, method, mtime, mdate, crc, csize, usize, nlen, xlen = struct.unpack_from("<IHHHHHIIIHH", zipbytes, 0)
    assert sig == 0x04034B50
    start = 30 + nlen + xlen
    return {"method": method, "flags": flags, "crc": crc, "csize": csize, "usize": usize, "start": start,
            "name": zipbytes[30:30 + nlen].decode("latin1")}


def main():
    out = {}
    src = open(os.path.join(REF, "Zip/ZipCorruptionHandling.cs"), encoding="utf-8-sig").read()
    good = base64.b64decode(const_string(src, "TestFileBadCDGoodCD64"))
    info = first_entry_payload(good)
    # sizes live in the zip64 extra / central directory for this fixture; the payload runs to the next signature
    end = good.index(b"PK\x01\x02")
    raw = good[info["start"]:end]
    out["inflate_ok"] = {"source": "Zip/ZipCorruptionHandling.cs:54-69 TestFileBadCDGoodCD64", "raw_hex": raw.hex(),
                         "entry": info}
    bad = base64.b64decode(const_string(src, "TestFileZeroCodeLength"))
    info2 = first_entry_payload(bad)
    raw2 = bad[info2["start"]:info2["start"] + info2["csize"]]  # csize is present in this local header (0x8D)
    out["inflate_zero_codelength"] = {"source": "Zip/ZipCorruptionHandling.cs:12-52 TestFileZeroCodeLength",
                                      "raw_hex": raw2.hex(), "entry": info2,
                                      "expect": "SharpZipBaseException (Encountered invalid codelength 0), must terminate"}
    out["checksum_kats"] = {
        "source": "Checksum/ChecksumTests.cs:24-37, 107-146",
        "adler32": [{"ascii": "123456789", "value": 0x091E01DE}],
        "crc32": [{"ascii": "123456789", "value": 0xCBF43926},
                  {"ascii": "123456789" * 4, "value": 0x3E29169C},
                  {"ascii": "456", "value": 0xB1A8C371, "note": "unaligned slice of 123456789 (offset 3, count 3)"},
                  {"ascii": "789123456789123456", "value": 0x31CA9A2E, "note": "offset 6 count 18 of 123456789 x4"}],
    }
    enc = open(os.path.join(REF, "Zip/ZipEncryptionHandling.cs"), encoding="utf-8-sig").read()
    m = re.search(r'const string TestFileWithEmptyPassword\s*=\s*@"([^"]*)"', enc)
    text = re.search(r'Is\.EqualTo\("(Lorem ipsum[^"]*)"\)', enc).group(1)
    out["aes_empty_password_zip"] = {"source": "Zip/ZipEncryptionHandling.cs:452-482 TestFileWithEmptyPassword / ZipFileAESReadWithEmptyPassword",
                                     "zip_base64": "".join(m.group(1).split()), "password": "", "entry": "test", "text": text}
    zf = open(os.path.join(REF, "Zip/ZipFileHandling.cs"), encoding="utf-8-sig").read()
    zf = zf[zf.index("public void ShouldReadAESBZip2ZipCreatedBy7Zip"):]  # (the plain bzip2 test in front of it uses the same names)
    text7 = re.search(r'const string originalText =\s*"([^"]*)"', zf).group(1)
    out["aes_bzip2_zip_by_7zip"] = {"source": "Zip/ZipFileHandling.cs:1707-1745 ShouldReadAESBZip2ZipCreatedBy7Zip",
                                    "zip_base64": const_string(zf, "bZip2CompressedZipCreatedBy7Zip"), "password": "password",
                                    "entry": "Hello.txt", "method": "bzip2", "text": text7}
    with open(os.path.join(HERE, "reference_fixtures.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote reference_fixtures.json:", {k: (len(v.get("raw_hex", "")) // 2 if isinstance(v, dict) else 0) for k, v in out.items()})


if __name__ == "__main__":
    main()
