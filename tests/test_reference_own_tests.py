"""The reference's OWN unit tests, unmodified, on the stand-in xarray (CPU,
build container only: /root/reference does not travel).

VERDICT r1 asked what validates ``weatherbenchx_b200.xarray_lite`` -- the
container the golden vectors were generated on -- against the reference's own
expectations.  Here the hot-path tests of the reference are imported from
/root/reference and run as they are:

  weatherbenchX/aggregation_test.py   AggregationTest (RMSE == 1, missing
      reduce dims, NaN / mask / skipna, weights x 4, bins add dims, DataTree
      and Dataset round trips; :69-270)
  weatherbenchX/weighting_test.py     WeightingTest.test_latitude_weights (:24-46)
  weatherbenchX/metrics/metrics_test.py  test_crps (8 parameterisations against
      the brute-force estimator, :603-660) and test_acc (:983-1006)

They exercise the stand-in's alignment, broadcasting by name, ``where``,
``mean(skipna)``, ``dot``, ``sel`` with slices, ``expand_dims`` and the
reductions -- with the reference's code on top and the reference's assertions
as the judge.
"""

import os
import sys
import unittest

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                'golden'))
import reference_runtime  # noqa: E402

pytestmark = pytest.mark.skipif(
    not reference_runtime.available(),
    reason='/root/reference is only present in the build container')


def _run(case_class, names=None):
  loader = unittest.TestLoader()
  if names is None:
    suite = loader.loadTestsFromTestCase(case_class)
  else:
    all_names = loader.getTestCaseNames(case_class)
    picked = [n for n in all_names if any(n.startswith(p) for p in names)]
    assert picked, (names, all_names)
    suite = unittest.TestSuite(case_class(n) for n in picked)
  result = unittest.TestResult()
  suite.run(result)
  problems = [f'{t}: {tb}' for t, tb in result.errors + result.failures]
  assert not problems, '\n'.join(problems)
  assert result.testsRun > 0
  return result.testsRun


@pytest.fixture(scope='module')
def reference():
  try:
    from absl.testing import absltest  # noqa: F401
  except ImportError:
    pytest.skip('absl.testing is not installed')
  return reference_runtime.install()


def test_reference_aggregation_tests(reference):
  from weatherbenchX import aggregation_test
  assert _run(aggregation_test.AggregationTest) == 7


def test_reference_weighting_test(reference):
  from weatherbenchX import weighting_test
  assert _run(weighting_test.WeightingTest) == 1


def test_reference_crps_and_acc_tests(reference):
  from weatherbenchX.metrics import metrics_test
  assert _run(metrics_test.MetricsTest, ['test_crps', 'test_acc']) >= 9
