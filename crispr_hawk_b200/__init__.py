"""crispr_hawk_b200 -- B200-native guide-discovery scan behind CRISPR-HAWK's own API.

    from crispr_hawk_b200 import install
    install()                      # then run `crisprhawk search ...` unchanged

Public mirror of the reference interface for this path:
`encode`, `encode_haplotypes` (encoder.py / crisprhawk.py:64), `PAM` (pam.py),
`search`, `pam_search` (search_guides.py). The compute runs in libhawkscan.so
(hand-written CUDA for sm_100a) through the C-ABI in include/hawkscan.h.
"""

from .annotation import annotate_table  # noqa: F401
from .encoder import encode, encode_haplotypes, encode_region  # noqa: F401
from .guide import Guide, guide_class  # noqa: F401
from .guide_table import GuideTable  # noqa: F401
from .haplotypes import Edit, EditHaplotype, build_phased  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .pam import PAM  # noqa: F401
from .search_guides import pam_search, search, search_table  # noqa: F401

__version__ = "0.1.0"
