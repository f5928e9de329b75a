"""CPU: the oracle restatement reproduces the fixtures that the REAL reference produced
(oracle/make_golden.py).  fp32, ||a-b||inf / ||b||inf <= 1e-5; bool masks and prediction counts exact."""
import pytest
import torch

import _cases as C


@pytest.mark.parametrize("name", C.golden_files("decoder"))
def test_decoder_golden(name):
    g = C.load_golden(name)
    C.assert_close_to_golden(C.oracle_decoder(g["case"]), g["out"])


@pytest.mark.parametrize("name", C.golden_files("maskhead"))
def test_maskhead_golden(name):
    g = C.load_golden(name)
    C.assert_close_to_golden(C.oracle_maskhead(g["case"]), g["out"])


@pytest.mark.parametrize("name", C.golden_files("model"))
def test_model_golden(name):
    g = C.load_golden(name)
    C.assert_close_to_golden(C.oracle_model(g["case"]), g["out"])


def test_golden_present():
    assert len(C.golden_files()) >= 10
