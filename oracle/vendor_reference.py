"""Recipe that makes the UNMODIFIED reference travel to the GPU box.

    python -m oracle.vendor_reference

Copies the Python sources of /root/reference/promonet (plus config/ and the
small assets/stats + assets/configs it reads at import) into oracle/_ref/, which
is git-ignored (the reference never enters this repository's history) but NOT
gpurun-ignored, so bench.py's CPU-baseline / `--impl reference` legs and the
PyTorch-eager-on-B200 bar can run the reference's own modules there.  Product
code never imports it; `__graft_entry__.build()` runs this when /root/reference
exists and leaves oracle/_ref/ alone otherwise.  Nothing is edited: a checksum
of every copied file is written next to the copy (MANIFEST.json).
"""
import hashlib
import json
import shutil
import sys
from pathlib import Path

SOURCE = Path('/root/reference')
TARGET = Path(__file__).resolve().parent / '_ref'
KEEP_ASSETS = ('stats', 'configs')


def vendor(source=SOURCE, target=TARGET):
    source, target = Path(source), Path(target)
    if not (source / 'promonet').is_dir():
        return None
    manifest = {}
    if target.exists():
        shutil.rmtree(target)
    for file in sorted(source.glob('promonet/**/*')):
        relative = file.relative_to(source)
        if not file.is_file() or '__pycache__' in relative.parts:
            continue
        if relative.parts[1] == 'assets' and (
                len(relative.parts) < 4 or relative.parts[2] not in KEEP_ASSETS):
            continue
        if relative.parts[1] != 'assets' and file.suffix != '.py':
            continue
        manifest[str(relative)] = file
    for file in sorted(source.glob('config/**/*.py')):
        manifest[str(file.relative_to(source))] = file
    digests = {}
    for relative, file in manifest.items():
        destination = target / relative
        destination.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(file, destination)
        digests[relative] = hashlib.sha256(file.read_bytes()).hexdigest()
    (target / 'MANIFEST.json').write_text(json.dumps(
        {'source': str(source), 'files': digests}, indent=1, sort_keys=True))
    return target


if __name__ == '__main__':
    result = vendor(*(sys.argv[1:2]))
    print(result if result else f'{SOURCE} not present: nothing copied')
