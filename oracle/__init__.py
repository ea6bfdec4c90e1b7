"""CPU oracle for the two alignment loss heads.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package imports this directory.  Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import it, and there only as the checker or as the
thing timed as "the reference's CPU path" — never as a code path of the product.

Parity status: PINNED.  The reference's own tests hold no golden value for this path, so the
oracle is pinned against outputs of the reference itself: tests/golden/make_golden.py imports the reference
modules from /root/reference, runs them on seeded inputs and stores inputs + outputs under
tests/golden/*.npz; tests/test_oracle.py checks both restatements below against those fixtures.

  closed_form.py  numpy float64 restatement (losses AND analytic gradients), one function per reference
                  function, each citing the file:line it follows.
  torch_port.py   the same arithmetic written with the torch ops the reference uses (matmul,
                  cross_entropy, log_softmax, gather, logsigmoid) so that bench.py can time "the reference's
                  CPU implementation" on a box where /root/reference does not exist.
  ref_loader.py   imports the real reference modules when /root/reference is present (this container only).
"""
