"""TEST INFRASTRUCTURE -- writes tests/golden/model_blocks.yaml: the `model:` blocks of the reference's
own experiment configs (expt/nytimes/{4_no_image,5_transformer_roberta,8_transformer_faces,
9_transformer_objects}/config.yaml), keys and values verbatim.  Run in the build container only."""
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('TT_REFERENCE_ROOT', '/root/reference')
NAMES = ('4_no_image', '5_transformer_roberta', '8_transformer_faces', '9_transformer_objects')

if __name__ == '__main__':
    out = {}
    for name in NAMES:
        with open(os.path.join(REF, 'expt', 'nytimes', name, 'config.yaml')) as f:
            out[name] = yaml.safe_load(f)['model']
    head = ("# `model:` blocks of the reference's expt/nytimes/<name>/config.yaml, verbatim (keys and values),\n"
            "# extracted by oracle/extract_model_blocks.py.  They are the constructor contract of SURVEY 8(b):\n"
            "# tests/test_abi_cpu.py instantiates each through tell_b200.registry unchanged.\n")
    with open(os.path.join(ROOT, 'tests', 'golden', 'model_blocks.yaml'), 'w') as f:
        f.write(head + yaml.safe_dump(out, sort_keys=False))
