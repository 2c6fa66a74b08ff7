"""The reference's known-answer cases (tests/golden_cases.py) run through the C ABI of the CUDA library."""
import pytest

from tests.golden_cases import golden_cases, run_golden_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c.__name__)
def test_reference_known_answers_on_gpu(product_fns, case):
    run_golden_case(case, product_fns)
