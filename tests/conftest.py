import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kosmos-x_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def tiny_cfgs():
    import kosmos_oracle as ko
    from kosmosx import KosmosConfig
    oc = ko.OracleConfig.tiny()
    kc = KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__})
    return oc, kc
