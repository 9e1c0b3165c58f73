"""Pins the CPU oracle against the golden values of the reference's own unit tests.

Every case names the reference test it restates (file:line under /root/reference/src/).
Comparison is exact f32 equality where the reference uses assert_eq!, and 5e-6 where it uses
assert_epsilon! (block_helpers.rs:30-40).
"""
import numpy as np
import pytest

from oracle import fw_oracle as fo

M31 = 0x7FFFFFFF
NOF = 0x80000000
ONE = 1065353216


def f32(x):
    return np.float32(x)


def eq(got, want):
    assert np.float32(got) == np.float32(want), (float(got), want)


def close(got, want, eps=5e-6):
    assert abs(float(got) - want) < eps, (float(got), want)


def nd(start, end):
    return (start << 16) + end


def bits(x):
    return int(np.float32(x).view(np.uint32))


# ----------------------------------------------------------------------------- hashing
def test_murmur3_known_answers():
    # parser.rs:490-727: "1 |A a" -> 2988156968, |B b -> 2422381320, |A b -> 3529656005, |A c -> 906509
    seed_a = fo.murmur3_32(b"A", 0)
    seed_b = fo.murmur3_32(b"B", 0)
    assert fo.murmur3_32(b"a", seed_a) == 2988156968
    assert fo.murmur3_32(b"b", seed_b) == 2422381320
    assert fo.murmur3_32(b"b", seed_a) == 3529656005
    assert fo.murmur3_32(b"c", seed_a) & M31 == 906509 & M31
    # parser.rs:1033-1046 multi-byte namespace
    assert fo.murmur3_32(b"a", fo.murmur3_32(b"AA", 0)) & M31 == 292540976 & M31
    # parser.rs:873-885 |B 3 -> 1775699190
    assert fo.murmur3_32(b"3", seed_b) & M31 == 1775699190 & M31


# ----------------------------------------------------------------------------- parser
def test_parser_vowpal():  # parser.rs:475-857 test_vowpal
    p = fo.Parser(["A", "B", "C"])
    a = 2988156968 & M31
    for line in ("1 |A a\n", "1 |A a \n", "1  |A a\n", "1 |A  a\n", "1 |A:1.0 a\n"):
        assert p.parse(line).tolist() == [6, 1, ONE, a, NOF, NOF], line
    assert p.parse("-1 |B b\n").tolist() == [6, 0, ONE, NOF, 2422381320 & M31, NOF]
    assert p.parse("1 |A a b\n").tolist() == [10, 1, ONE, nd(6, 10) | NOF, NOF, NOF, a, ONE, 3529656005 & M31, ONE]
    assert p.parse("-1 |A a |B b\n").tolist() == [6, 0, ONE, a, 2422381320 & M31, NOF]
    assert p.parse("-1 |A a  |B b\n").tolist() == [6, 0, ONE, a, 2422381320 & M31, NOF]
    with pytest.raises(ValueError, match="Feature name was not predeclared in vw_namespace_map.csv: UNDECLARED_NAMESPACE"):
        p.parse("1 |UNDECLARED_NAMESPACE a\n")
    with pytest.raises(ValueError, match="Failed parsing namespace weight: not_a_parsable_number"):
        p.parse("1 |A:not_a_parsable_number a\n")
    with pytest.raises(ValueError, match="Failed parsing namespace weight: 1:1"):
        p.parse("1 |A:1:1 a\n")
    assert p.parse("1 |A:2.0 a\n").tolist() == [8, 1, ONE, nd(6, 8) | NOF, NOF, NOF, a, bits(2.0)]
    assert p.parse("1 |A a:2.0\n").tolist() == [8, 1, ONE, nd(6, 8) | NOF, NOF, NOF, a, bits(2.0)]
    assert p.parse("1 |A a:2.0 b:3.0\n").tolist() == [10, 1, ONE, nd(6, 10) | NOF, NOF, NOF, a, bits(2.0),
                                                       3529656005 & M31, bits(3.0)]
    assert p.parse("1 |A:3 a:2.0\n").tolist() == [8, 1, ONE, nd(6, 8) | NOF, NOF, NOF, a, bits(6.0)]
    with pytest.raises(ValueError, match="Failed parsing feature weight: 2x0"):
        p.parse("1 |A a:2x0\n")
    assert p.parse("1 |A a b:2.0 c:3.0\n").tolist() == [12, 1, ONE, nd(6, 12) | NOF, NOF, NOF, a, bits(1.0),
                                                         3529656005 & M31, bits(2.0), 906509 & M31, bits(3.0)]
    assert p.parse("|A a\n").tolist() == [6, 0xFF, ONE, a, NOF, NOF]
    assert p.parse("").tolist() == []
    with pytest.raises(fo.FlushCommand):
        p.parse("flush")
    with pytest.raises(ValueError, match="Cannot parse an example"):
        p.parse("$1")
    with pytest.raises(ValueError, match="Example importance cannot be negative: -0.1! "):
        p.parse("1 -0.1 |A a\n")
    with pytest.raises(ValueError, match="Failed parsing example importance: fdsa"):
        p.parse("1 fdsa |A a\n")
    assert p.parse("1 0.1 |A a\n").tolist() == [6, 1, bits(0.1), a, NOF, NOF]
    assert p.parse("1  0.1  |A  a \n").tolist() == [6, 1, bits(0.1), a, NOF, NOF]
    with pytest.raises(fo.HogwildLoadCommand):
        p.parse("hogwild_load /path/to/filename")
    with pytest.raises(fo.HogwildLoadCommand):
        p.parse("hogwild_load   /path/to/filename  ")
    with pytest.raises(ValueError, match="Cannot parse an example"):
        p.parse("hogwild_load")
    with pytest.raises(ValueError, match="Cannot parse an example"):
        p.parse("hogwild_load ")


def test_parser_float_namespaces():  # parser.rs:860-1015
    p = fo.Parser(["A", "B", "C"])
    assert p.parse("-1 |B 3\n").tolist() == [6, 0, ONE, NOF, 1775699190 & M31, NOF]
    p = fo.Parser(["A", "B", "C"], ns_is_f32=[0, 1, 0])
    assert p.parse("-1 |B 3\n").tolist() == [8, 0, ONE, NOF, nd(6, 8) | NOF, NOF, 1775699190 & M31, bits(3.0)]
    want = [10, 0, ONE, NOF, nd(6, 10) | NOF, NOF, 1775699190 & M31, bits(3.0), 382082293 & M31, bits(4.0)]
    assert p.parse("-1 |B 3 4\n").tolist() == want
    with pytest.raises(ValueError, match=r"Failed parsing feature value to float \(for float namespace\): not_a_number"):
        p.parse("-1 |B not_a_number\n")
    assert p.parse("-1 |B 3 4\n").tolist() == want
    with pytest.raises(ValueError, match="Namespaces that are f32 can not have weight attached"):
        p.parse("-1 |B 3:3\n")
    with pytest.raises(ValueError, match="Namespaces that are f32 can not have weight attached"):
        p.parse("-1 |B:3 3\n")
    p = fo.Parser(["A", "B", "C"], ns_is_f32=[0, 1, 0], namespace_skip_prefix=1)
    assert p.parse("-1 |B B3\n").tolist() == [8, 0, ONE, NOF, nd(6, 8) | NOF, NOF, 1416737454 & M31, bits(3.0)]
    nan_bits = 0x7FC00000  # f32::NAN.to_bits()
    r = p.parse("-1 |B B\n").tolist()
    assert r[:7] == [8, 0, ONE, NOF, nd(6, 8) | NOF, NOF, 25602353 & M31] and r[7] == nan_bits
    r = p.parse("-1 |B BNONE\n").tolist()
    assert r[:7] == [8, 0, ONE, NOF, nd(6, 8) | NOF, NOF, 1846432377 & M31] and r[7] == nan_bits


def test_parser_multibyte_namespaces():  # parser.rs:1018-1061
    p = fo.Parser(["AA", "BB", "CC"])
    assert p.parse("1 |AA a\n").tolist() == [6, 1, ONE, 292540976 & M31, NOF, NOF]
    assert p.parse("1 |AA:3 a:2.0\n").tolist() == [8, 1, ONE, nd(6, 8) | NOF, NOF, NOF, 292540976 & M31, bits(6.0)]


# ----------------------------------------------------------------------------- translate
def hdr(v):
    return [100, 1, ONE] + v


def test_translate_constant():  # feature_buffer.rs:375-395
    s = fo.Spec(1, [[0]], add_constant=True)
    _, _, lr, _ = s.translate(hdr([NOF]))
    assert lr == [(116060, 1.0, 1)]


def test_translate_single_once():  # feature_buffer.rs:397-447
    s = fo.Spec(1, [[0]], add_constant=False)
    assert s.translate(hdr([NOF]))[2] == []
    assert s.translate(hdr([0xFEA]))[2] == [(0xFEA, 1.0, 0)]
    rb = hdr([NOF | nd(4, 8), 0xFEA, ONE, 0xFEB, ONE])
    assert s.translate(rb)[2] == [(0xFEA, 1.0, 0), (0xFEB, 1.0, 0)]


def test_translate_single_twice():  # feature_buffer.rs:449-500
    s = fo.Spec(2, [[0], [1]], add_constant=False)
    assert s.translate(hdr([NOF, NOF]))[2] == []
    assert s.translate(hdr([0xFEA, NOF]))[2] == [(0xFEA, 1.0, 0)]
    assert s.translate(hdr([0xFEA, 0xFEB]))[2] == [(0xFEA, 1.0, 0), (0xFEB, 1.0, 1)]


def test_translate_double_vowpal():  # feature_buffer.rs:502-535 interaction hash 208368
    s = fo.Spec(2, [[0, 1]], add_constant=False)
    assert s.translate(hdr([NOF, NOF]))[2] == []
    assert s.translate(hdr([123456789, NOF]))[2] == []
    assert s.translate(hdr([2988156968 & M31, 2422381320 & M31, NOF]))[2] == [(208368, 1.0, 0)]


def test_translate_single_with_weight():  # feature_buffer.rs:537-560
    s = fo.Spec(1, [[0]], combo_weights=[2.0], add_constant=False)
    assert s.translate(hdr([0xFEA]))[2] == [(0xFEA, 2.0, 0)]


def test_translate_ffm():  # feature_buffer.rs:562-742
    s = fo.Spec(1, [], add_constant=False, fields=[[]], ffm_k=1)
    assert s.translate(hdr([0xFEA]))[3] == []
    s = fo.Spec(1, [], add_constant=False, fields=[[0]], ffm_k=1)
    assert s.translate(hdr([0xFEA]))[3] == [(0xFEA, 1.0, 0)]
    s = fo.Spec(2, [], add_constant=False, fields=[[0], [0, 1]], ffm_k=1)
    rb = hdr([NOF | nd(5, 9), 0xFEC, 0xFEA, bits(2.0), 0xFEB, bits(3.0)])
    assert s.translate(rb)[3] == [(0xFEA, 2.0, 0), (0xFEB, 3.0, 0), (0xFEA, 2.0, 1), (0xFEB, 3.0, 1), (0xFEC, 1.0, 1)]
    rb = hdr([NOF | nd(5, 9), 0x1, 0xFFF, bits(2.0), 0xFEB, bits(3.0)])
    s = fo.Spec(2, [], add_constant=False, fields=[[0], [0, 1], [1]], ffm_k=1)
    assert s.translate(rb)[3] == [(0xFFF, 2.0, 0), (0xFEB, 3.0, 0), (0xFFF, 2.0, 1), (0xFEB, 3.0, 1), (0x1, 1.0, 1),
                                  (0x1, 1.0, 2)]
    s = fo.Spec(2, [], add_constant=False, fields=[[0], [0, 1], [1]], ffm_k=3)
    assert s.translate(rb)[3] == [(0xFFC, 2.0, 0), (0xFE8, 3.0, 0), (0xFFC, 2.0, 3), (0xFE8, 3.0, 3), (0x0, 1.0, 3),
                                  (0x0, 1.0, 6)]


def test_translate_f32_namespace():  # feature_buffer.rs:759-796
    s = fo.Spec(3, [[1]], add_constant=False, ns_is_f32=[0, 1, 0])
    rb = hdr([NOF, nd(6, 10) | NOF, NOF, 0xFFC & M31, bits(3.0), 0xFFA & M31, bits(4.0)])
    assert s.translate(rb)[2] == [(0xFFC, 1.0, 0), (0xFFA, 1.0, 0)]


# ----------------------------------------------------------------------------- optimizer
def test_optimizer_goldens():  # optimizer.rs:170-226
    L = fo.lib()
    import ctypes as C

    def upd(kind, lr, pt, init, g, acc):
        lut = fo.lut_build(lr, pt, init)
        a = C.c_float(acc)
        u = L.fwo_opt_update(kind, lr, -pt, lut.ctypes.data_as(C.POINTER(C.c_float)), g, C.byref(a))
        return u, a.value

    # test_sgd: lr .15 -> update(0.1) = 0.1*0.15 ; (optimizer.rs:172-180)
    u, _ = upd(fo.OPT_SGD, 0.15, 0.4, 0.0, 0.1, 0.0)
    eq(u, f32(0.1) * f32(0.15))


def test_optimizer_flex_and_lut():  # optimizer.rs:181-226, 229-268
    import ctypes as C

    L = fo.lib()
    lr, pt, init = 0.15, 0.4, 0.0
    lut = fo.lut_build(lr, pt, init)
    lutp = lut.ctypes.data_as(C.POINTER(C.c_float))

    def upd(kind, g, acc):
        a = C.c_float(acc)
        u = L.fwo_opt_update(kind, lr, -pt, lutp, g, C.byref(a))
        return u, a.value

    # test_adagradflex (optimizer.rs:181-203)
    u, a = upd(fo.OPT_ADAGRAD_FLEX, 0.1, 0.9)
    eq(u, 0.015576674)
    eq(a, f32(0.9) + f32(0.1) * f32(0.1))
    u, a = upd(fo.OPT_ADAGRAD_FLEX, 0.1, 0.0)
    eq(u, 0.09464361)
    eq(a, f32(0.1) * f32(0.1))
    u, a = upd(fo.OPT_ADAGRAD_FLEX, 0.0, 0.0)
    eq(a, 0.0)
    # test_adagradlut (optimizer.rs:205-226)
    u, a = upd(fo.OPT_ADAGRAD_LUT, 0.1, 0.9)
    eq(u, 0.015607622)
    eq(a, f32(0.9) + f32(0.1) * f32(0.1))
    u, a = upd(fo.OPT_ADAGRAD_LUT, 0.1, 0.0)
    eq(u, 0.09375872)
    eq(a, f32(0.1) * f32(0.1))
    u, a = upd(fo.OPT_ADAGRAD_LUT, 0.0, 0.0)
    eq(u, 0.0)
    eq(a, 0.0)
    # test_adagradlut_comparison (optimizer.rs:229-268): relative error < 5 %
    for g in [-1.0, -0.9, -0.1, -0.00001, 0.0, 0.00001, 0.1, 0.5, 0.9, 1.0]:
        for acc in [0.0000000001, 0.00001, 0.1, 0.5, 1.1, 2.0, 20.0, 200.0, 2000.0, 200000.0, 2000000.0]:
            pf, _ = upd(fo.OPT_ADAGRAD_FLEX, g, acc)
            pl, _ = upd(fo.OPT_ADAGRAD_LUT, g, acc)
            err = abs(pf - pl)
            rel = err / abs(pl) if pl != 0.0 else err
            assert rel < 0.05, (g, acc, pf, pl)


# ----------------------------------------------------------------------------- LR regressor goldens
def lr_fb(feats, importance=1.0):
    return fo.feature_buffer(lr=feats, label=0.0, importance=importance)


def test_learning_turned_off():  # regressor.rs:556-594
    r = fo.Regressor(optimizer=fo.OPT_ADAGRAD_LUT)
    eq(r.learn(lr_fb([]), False), 0.5)
    eq(r.learn(lr_fb([(1, 1.0, 0)]), False), 0.5)
    eq(r.learn(lr_fb([(1, 1.0, 0), (2, 1.0, 0)]), False), 0.5)


def test_power_t_zero():  # regressor.rs:597-626
    for opt in (fo.OPT_ADAGRAD_FLEX, fo.OPT_ADAGRAD_LUT, fo.OPT_SGD):
        r = fo.Regressor(learning_rate=0.1, power_t=0.0, optimizer=opt)
        fb = lr_fb([(1, 1.0, 0)])
        eq(r.learn(fb), 0.5)
        eq(r.learn(fb), 0.48750263)
        eq(r.learn(fb), 0.47533244)


def test_double_same_feature():  # regressor.rs:629-656
    r = fo.Regressor(learning_rate=0.1, power_t=0.0, optimizer=fo.OPT_ADAGRAD_LUT)
    fb = lr_fb([(1, 1.0, 0), (1, 2.0, 0)])
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.38936076)
    eq(r.learn(fb), 0.30993468)


def test_power_t_half():  # regressor.rs:659-704
    r = fo.Regressor(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=fo.OPT_ADAGRAD_FLEX)
    fb = lr_fb([(1, 1.0, 0)])
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.4750208)
    eq(r.learn(fb), 0.45788094)


def test_power_t_half_fastmath():  # regressor.rs:707-748 (LUT, 11 bits)
    r = fo.Regressor(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=fo.OPT_ADAGRAD_LUT)
    fb = lr_fb([(1, 1.0, 0)])
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.475734)


def test_power_t_half_two_features():  # regressor.rs:751-812
    r = fo.Regressor(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=fo.OPT_ADAGRAD_FLEX)
    fb2 = lr_fb([(1, 1.0, 0), (2, 1.0, 0)])
    eq(r.learn(fb2), 0.5)
    eq(r.learn(fb2), 0.45016602)
    eq(r.learn(lr_fb([(1, 1.0, 0)])), 0.45836908)


def test_non_one_weight():  # regressor.rs:815-861
    r = fo.Regressor(learning_rate=0.1, power_t=0.0, optimizer=fo.OPT_ADAGRAD_LUT)
    fb = lr_fb([(1, 2.0, 0)])
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.45016602)
    eq(r.learn(fb), 0.40611085)


def test_example_importance():  # regressor.rs:864-884
    r = fo.Regressor(learning_rate=0.1, power_t=0.0, optimizer=fo.OPT_ADAGRAD_LUT)
    fb = lr_fb([(1, 1.0, 0)], importance=0.5)
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.49375027)
    eq(r.learn(fb), 0.4875807)


def test_save_load_and_test_mode_lr_values():  # persistence.rs:251-313 (arithmetic part)
    r = fo.Regressor(learning_rate=0.1, power_t=0.5, init_acc_gradient=0.0, optimizer=fo.OPT_ADAGRAD_FLEX)
    fb = lr_fb([(1, 1.0, 0), (2, 1.0, 0)])
    eq(r.learn(fb), 0.5)
    eq(r.learn(fb), 0.45016602)
    eq(r.learn(fb, False), 0.41731137)
    eq(r.predict(fb), 0.41731137)


# ----------------------------------------------------------------------------- FFM block goldens
def ffm_block(k, F, opt, **kw):
    """new_ffm_block -> new_logloss_block, weights forced to 1.0 (block_ffm.rs:1229-1236)."""
    args = dict(learning_rate=0.1, ffm_learning_rate=0.1, power_t=0.0, ffm_power_t=0.0, ffm_k=k,
                ffm_bit_precision=18, ffm_num_fields=F, optimizer=opt, graph=fo.GRAPH_FFM_BLOCK_ONLY)
    args.update(kw)
    r = fo.Regressor(**args)
    return r


def ones(r):
    r.ffm_weights[:] = 1.0
    r.ffm_acc[:] = r.desc.ffm_init_acc_gradient if r.desc.optimizer == fo.OPT_ADAGRAD_FLEX else 0.0
    return r


def ffm_fb(feats):
    return fo.feature_buffer(ffm=feats, label=0.0)


def test_ffm_k1():  # block_ffm.rs:1239-1323
    r = ffm_block(1, 2, fo.OPT_ADAGRAD_LUT)
    fb = ffm_fb([(1, 1.0, 0)])
    close(r.predict(fb), 0.5)
    close(r.forward_backward(fb, True), 0.5)
    r = ones(ffm_block(1, 2, fo.OPT_ADAGRAD_FLEX))
    fb = ffm_fb([(1, 1.0, 0), (100, 1.0, 1)])
    close(r.predict(fb), 0.7310586)
    eq(r.forward_backward(fb, True), 0.7310586)
    close(r.predict(fb), 0.7024794)
    eq(r.forward_backward(fb, True), 0.7024794)
    r = ones(ffm_block(1, 2, fo.OPT_ADAGRAD_LUT))
    fb = ffm_fb([(1, 2.0, 0), (100, 2.0, 1)])
    eq(r.predict(fb), 0.98201376)
    eq(r.forward_backward(fb, True), 0.98201376)
    eq(r.predict(fb), 0.81377685)
    eq(r.forward_backward(fb, True), 0.81377685)


def test_ffm_k4():  # block_ffm.rs:1450-1532
    r = ffm_block(4, 2, fo.OPT_ADAGRAD_LUT)
    fb = ffm_fb([(1, 1.0, 0)])
    for _ in range(2):
        eq(r.predict(fb), 0.5)
        eq(r.forward_backward(fb, True), 0.5)
    r = ones(ffm_block(4, 2, fo.OPT_ADAGRAD_FLEX))
    fb = ffm_fb([(1, 1.0, 0), (100, 1.0, 4)])
    eq(r.predict(fb), 0.98201376)
    eq(r.forward_backward(fb, True), 0.98201376)
    eq(r.predict(fb), 0.96277946)
    eq(r.forward_backward(fb, True), 0.96277946)
    r = ones(ffm_block(4, 2, fo.OPT_ADAGRAD_LUT))
    fb = ffm_fb([(1, 2.0, 0), (100, 2.0, 4)])
    eq(r.predict(fb), 0.9999999)
    eq(r.forward_backward(fb, True), 0.9999999)
    eq(r.predict(fb), 0.99685884)
    eq(r.forward_backward(fb, True), 0.99685884)


def test_ffm_multivalue():  # block_ffm.rs:1658-1701
    r = ones(ffm_block(1, 2, fo.OPT_ADAGRAD_LUT))
    fb = ffm_fb([(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 1)])
    close(r.predict(fb), 0.9933072)
    eq(r.forward_backward(fb, True), 0.9933072)
    close(r.predict(fb), 0.9395168)
    eq(r.forward_backward(fb, False), 0.9395168)
    close(r.predict(fb), 0.9395168)
    eq(r.forward_backward(fb, False), 0.9395168)


def test_ffm_multivalue_k4_nonzero_powert():  # block_ffm.rs:1778-1817 (default lr .5 / power_t .5, LUT)
    r = ones(ffm_block(4, 2, fo.OPT_ADAGRAD_LUT, learning_rate=0.5, ffm_learning_rate=0.5, power_t=0.5,
                       ffm_power_t=0.5))
    fb = ffm_fb([(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 4)])
    eq(r.predict(fb), 1.0)
    eq(r.forward_backward(fb, True), 1.0)
    eq(r.predict(fb), 0.9949837)
    eq(r.forward_backward(fb, False), 0.9949837)
    eq(r.forward_backward(fb, False), 0.9949837)


def test_ffm_missing_field():  # block_ffm.rs:1882-1944
    r = ones(ffm_block(1, 3, fo.OPT_ADAGRAD_FLEX))
    fb = ffm_fb([(1, 1.0, 0), (5, 1.0, 1), (100, 1.0, 2)])
    close(r.predict(fb), 0.95257413)
    eq(r.forward_backward(fb, False), 0.95257413)
    fb = ffm_fb([(5, 1.0, 1)])
    eq(r.predict(fb), 0.5)
    # The reference's last assertion (slearn2 == 0.62245935, :1943) is stale: the current training
    # path subtracts the self-interaction (block_ffm.rs:236-244) and yields 0.5, same as spredict2
    # one line above it.  SURVEY.md section 4 documents this; the oracle follows the code.
    eq(r.forward_backward(fb, True), 0.5)


# ----------------------------------------------------------------------------- full regressor (LR + FFM + triangle)
def full_reg(**kw):
    args = dict(learning_rate=0.1, power_t=0.0, bit_precision=18, ffm_k=1, ffm_bit_precision=18, ffm_power_t=0.0,
                ffm_learning_rate=0.1, ffm_num_fields=2, optimizer=fo.OPT_ADAGRAD_FLEX, num_combos=1)
    args.update(kw)
    r = fo.Regressor(**args)
    r.ffm_weights[:] = 1.0
    r.ffm_acc[:] = r.desc.ffm_init_acc_gradient
    return r


def test_save_load_and_test_mode_ffm_values():  # persistence.rs:342-421 (arithmetic part)
    r = full_reg()
    fb = ffm_fb([(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 1)])
    eq(r.learn(fb, True), 0.9933072)
    close(r.learn(fb, False), 0.9395168)
    close(r.predict(fb), 0.9395168)


def test_hogwild_load_values():  # persistence.rs:437-555 (arithmetic part: LR + FFM mixed)
    r1, r2 = full_reg(), full_reg()
    fb1 = fo.feature_buffer(lr=[(52, 0.5, 0), (2, 1.0, 0)], ffm=[(1, 0.5, 0), (3000, 1.0, 0), (101, 2.0, 1)])
    fb2 = fo.feature_buffer(lr=[(1, 1.0, 0), (2, 1.0, 0)], ffm=[(1, 1.0, 0), (3000, 1.0, 0), (100, 2.0, 1)])
    eq(r1.learn(fb1, True), 0.97068775)
    eq(r1.learn(fb1, False), 0.8922257)
    eq(r1.predict(fb1), 0.8922257)
    eq(r2.learn(fb2, True), 0.9933072)
    eq(r2.learn(fb2, False), 0.92719215)
    eq(r2.predict(fb2), 0.92719215)
    eq(r2.learn(fb1, False), 0.93763095)
    eq(r2.predict(fb1), 0.93763095)
    eq(r1.learn(fb2, False), 0.98559695)
    eq(r1.predict(fb2), 0.98559695)


# ----------------------------------------------------------------------------- blocks
def test_triangle_and_neuron_blocks():
    # block_neural.rs:507-581: one neuron, init One, inputs [2.0] -> 2.0; after one update with
    # lr 0.1 (SGD-equivalent) it yields 1.5.  Exercised through the head with a 1-combo LR input.
    r = fo.Regressor(learning_rate=0.1, power_t=0.0, nn_learning_rate=0.1, nn_power_t=0.0, optimizer=fo.OPT_ADAGRAD_FLEX,
                     nn_layers=[{"width": 2, "activation": "relu", "init": "one"}])
    assert r.nn_layer_count == 2
    assert r.nn_weights(0).shape[0] == (1 + 1) * 2 and r.nn_weights(1).shape[0] == (2 + 1 + 1)
    p = r.learn(lr_fb([(1, 1.0, 0)]), True)
    eq(p, 0.5)  # LR weights start at zero -> x = 0 -> everything 0


def test_neuron_layer_goldens():
    # block_neural.rs:507-537 test_simple: const input [2.0], one neuron init One, SGD nn lr 0.1, observe gradient 1.0:
    # 2.0, then 1.5 (w = 1 - 0.1*2 = 0.8, bias = -0.1)
    outs, d_in = fo.neuron_layer_test(fo.OPT_SGD, 0.1, 0.0, 0.0, 1, 1, fo.NN_INIT_ONE, False, [2.0], [1.0], 2)
    eq(outs[0, 0], 2.0)
    eq(outs[1, 0], 1.5)
    # block_neural.rs:539-581 test_two_neurons: both neurons see 2.0, then 1.5
    outs, _ = fo.neuron_layer_test(fo.OPT_SGD, 0.1, 0.0, 0.0, 1, 2, fo.NN_INIT_ONE, False, [2.0], [1.0, 1.0], 2)
    assert outs[0].tolist() == [2.0, 2.0]
    eq(outs[1, 0], 1.5)
    eq(outs[1, 1], 1.5)
    # the input gradient uses the PRE-update weights (block_neural.rs:283-284): first step 1.0 * 1.0 per neuron
    _, d_in = fo.neuron_layer_test(fo.OPT_SGD, 0.1, 0.0, 0.0, 1, 2, fo.NN_INIT_ONE, False, [2.0], [1.0, 1.0], 1)
    eq(d_in[0], 2.0)
    # block_relu.rs:156-173: relu passes 2.0 and "doesn't learn" (no weights of its own); a clamped output blocks the gradient
    outs, d_in = fo.neuron_layer_test(fo.OPT_SGD, 0.1, 0.0, 0.0, 1, 1, fo.NN_INIT_ONE, True, [-2.0], [1.0], 2)
    assert outs[:, 0].tolist() == [0.0, 0.0] and d_in[0] == 0.0


def test_head_matches_numpy_restatement():
    """The oracle's whole head (copy -> layers -> join -> neuron -> sigmoid) against an independent numpy float32
    restatement of regressor.rs:191-320 on random weights: forward value and every updated parameter."""
    rng = np.random.default_rng(4)
    r = fo.Regressor(learning_rate=0.1, power_t=0.0, nn_learning_rate=0.1, nn_power_t=0.0, optimizer=fo.OPT_SGD, bit_precision=8,
                     num_combos=3, nn_layers=[{"width": 4, "activation": "relu"}, {"width": 3, "activation": "none"}])
    r.lr_table[:, 0] = rng.normal(0, 0.5, r.lr_table.shape[0]).astype(np.float32)
    for l in range(r.nn_layer_count):
        r.nn_weights(l)[:] = rng.normal(0, 0.5, r.nn_weights(l).shape[0]).astype(np.float32)
    fb = lr_fb([(1, 1.0, 0), (2, 2.0, 1), (3, 0.5, 2)])
    fb.label = 1.0
    f32 = np.float32
    x = np.array([r.lr_table[1, 0] * f32(1.0), r.lr_table[2, 0] * f32(2.0), r.lr_table[3, 0] * f32(0.5)], f32)
    W = [r.nn_weights(l).copy() for l in range(3)]
    def layer(w, n_in, n_out, inp):
        return (w[:n_in * n_out].reshape(n_out, n_in).astype(np.float64) @ inp.astype(np.float64)).astype(f32) + w[n_in * n_out:]
    z0 = layer(W[0], 3, 4, x); h0 = np.where(z0 < 0, f32(0), z0)
    h1 = layer(W[1], 4, 3, h0)
    y = layer(W[2], 6, 1, np.concatenate([h1, x]))[0]
    p_want = 1.0 / (1.0 + np.exp(-np.float64(y)))
    p = r.learn(fb, True)
    assert abs(p - p_want) < 2e-6, (p, p_want)
    g = f32(-(1.0 - p))
    # SGD lr 0.1: final neuron w -= 0.1 * g * in ; hidden layers through the pre-update weights
    in2 = np.concatenate([h1, x])
    w2 = W[2].copy(); w2[:6] -= f32(0.1) * g * in2; w2[6] -= f32(0.1) * g
    np.testing.assert_allclose(r.nn_weights(2), w2, rtol=0, atol=1e-6)
    d_h1 = W[2][:3] * g
    w1 = W[1].copy(); w1[:12] -= (f32(0.1) * np.outer(d_h1, h0)).reshape(-1); w1[12:] -= f32(0.1) * d_h1
    np.testing.assert_allclose(r.nn_weights(1), w1, rtol=0, atol=1e-6)
    d_h0 = (W[1][:12].reshape(3, 4).T @ d_h1) * (z0 >= 0)
    w0 = W[0].copy(); w0[:12] -= (f32(0.1) * np.outer(d_h0, x)).reshape(-1); w0[12:] -= f32(0.1) * d_h0
    np.testing.assert_allclose(r.nn_weights(0), w0, rtol=0, atol=1e-6)
    d_x = W[0][:12].reshape(4, 3).T @ d_h0 + W[2][3:6] * g   # BlockCopy backward adds both paths (block_misc.rs:452-473)
    want_lr = [r_w - f32(0.1) * d * v for r_w, d, v in zip([x[0] / f32(1.0), x[1] / f32(2.0), x[2] / f32(0.5)], d_x, [1.0, 2.0, 0.5])]
    np.testing.assert_allclose([r.lr_table[1, 0], r.lr_table[2, 0], r.lr_table[3, 0]], want_lr, rtol=0, atol=1e-6)


def test_merand48_range_and_determinism():
    # parity unpinned (no reference test observes it) -- sanity only
    xs = [fo.merand48(i) for i in range(1000)]
    assert all(0.0 <= x < 1.0 for x in xs)
    assert len(set(xs)) > 990
    assert abs(np.mean(xs) - 0.5) < 0.05


def test_head_wave_tool_with_one_example_in_flight_is_the_sequential_learner():
    """fwo_learn_records_head_wave (the batched-head semantics the device is compared with) collapses to fwo_learn when the
    sub-batch is one example: same predictions and the same tables, bit for bit -- which pins the tool's dense backward,
    triangle backward and sparse update to the restated reference arithmetic."""
    from fwumious_wabbit_b200 import synth
    from tests import util

    n_ns, k = 6, 4
    w = synth.Workload("c5t", synth._mi(n_ns, ffm_k=k, ffm_bits=12, bits=12, lr=0.05, ffm_lr=0.02, ffm_init_acc=0.1),
                       synth.NS_LETTERS[:n_ns], [30] * n_ns, "tiny head model")
    w.mi.nn_layers = [{"width": "8", "activation": "relu"}, {"width": "5", "activation": "relu"}]
    w.mi.nn_learning_rate, w.mi.nn_power_t, w.mi.nn_init_acc_gradient = 0.02, 0.5, 0.1
    n = 400
    recs = w.records(n)
    recs[::7, 2] = np.float32(0.0).view(np.uint32)      # importance 0: scored, not learned
    recs[5::11, 3 + 2] = 0x80000000                      # an absent namespace
    spec = util.oracle_spec(w.mi)
    rec_off = np.arange(n + 1, dtype=np.uint64) * w.record_len
    a, b = util.oracle_regressor(w.mi), util.oracle_regressor(w.mi)
    _, want = a.hogwild(spec, recs.reshape(-1), rec_off, 1, want_preds=True)
    got = b.learn_head_wave(spec, recs.reshape(-1), rec_off, 1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(a.ffm_weights.view(np.uint32), b.ffm_weights.view(np.uint32))
    assert np.array_equal(a.lr_table.view(np.uint32), b.lr_table.view(np.uint32))
    for l in range(a.nn_layer_count):
        assert np.array_equal(a.nn_weights(l).view(np.uint32), b.nn_weights(l).view(np.uint32))
        assert np.array_equal(a.nn_acc(l).view(np.uint32), b.nn_acc(l).view(np.uint32))
    # a real sub-batch differs (staleness) but stays close and still learns
    c = util.oracle_regressor(w.mi)
    p64 = c.learn_head_wave(spec, recs.reshape(-1), rec_off, 64)
    assert np.all(np.isfinite(p64)) and np.max(np.abs(p64[:64] - want[:1])) < 1.0
