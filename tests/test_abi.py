"""The C-ABI library loads and exports every symbol include/rxn_b200.h declares; the ctypes
mirror of RxnTablesDesc has the compiled layout.  No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from pflotran_b200 import abi, reactive_transport as rt
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'rxn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rxn_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(rt.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(rt.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), 'librxn_b200.so does not export %s' % n


def test_desc_layout_matches_compiler():
    assert pyoracle.lib().orc_desc_size() == C.sizeof(abi.RxnTablesDesc)


def test_field_enum_matches_header():
    text = open(os.path.join(ROOT, 'include', 'rxn_b200.h')).read()
    body = text[text.index('typedef enum RxnField'):text.index('} RxnField;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'RXN_F_([A-Z0-9_]+)', body)
    names = [n for n in names if n != 'COUNT']
    assert names == abi.FIELDS


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product refuses to run (there is no CPU fallback)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip('GPU present')
    from pflotran_b200 import synth
    w = synth.Workload('calcite')
    with pytest.raises(rt.RxnError) as e:
        rt.Reaction(w.tables)
    assert e.value.status in (abi.RXN_ERR_NO_DEVICE, abi.RXN_ERR_CUDA)


def test_product_does_not_import_oracle():
    """Nothing under pflotran_b200/ or include/ references oracle/ or the test emulation."""
    bad = []
    for base in ('pflotran_b200', 'include'):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if 'build' in dp or '__pycache__' in dp:
                continue
            for f in fn:
                if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp', 'Makefile')):
                    t = open(os.path.join(dp, f)).read()
                    if re.search(r'(from|import)\s+oracle|oracle/|liborc|libemul|tests/emul', t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_fortran_shim_binds_exported_symbols_in_struct_order():
    """fortran/rxn_b200_shim.F90 (not compilable here: no Fortran compiler) names only exported
    symbols, and its bind(C) mirror of RxnTablesDesc lists the fields in the order of the C struct."""
    src = open(os.path.join(ROOT, 'fortran', 'rxn_b200_shim.F90')).read()
    bound = re.findall(r"bind\(C,\s*name='(rxn_[a-z_0-9]+)'\)", src)
    assert len(bound) >= 18
    L = C.CDLL(rt.LIB_PATH)
    for n in bound:
        assert hasattr(L, n), 'shim binds %s which librxn_b200.so does not export' % n
    body = src[src.index('type, bind(C), public :: rxn_tables_desc_type'):src.index('end type rxn_tables_desc_type')]
    body = re.sub(r'!.*', '', body).replace('&\n', ' ')
    f_fields = []
    for line in body.splitlines()[1:]:
        if '::' in line:
            f_fields += [x.strip() for x in line.split('::')[1].split(',') if x.strip()]
    c_fields = [n for n, _ in abi.RxnTablesDesc._fields_]
    assert [f.lower() for f in f_fields] == [c.lower() for c in c_fields]
    # field enum values
    for i, name in enumerate(abi.FIELDS):
        m = re.search(r'RXN_F_%s\s*=\s*(\d+)' % name, src)
        assert m and int(m.group(1)) == i, name
