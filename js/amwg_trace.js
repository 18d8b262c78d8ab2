// amwg_trace.js -- record a log_post(state, data) closure and lower it to the program the CUDA sampler runs (include/amwg.h).
//
// The JavaScript counterpart of bayes.js_b200/tracer.py: same expression nodes, same lowering, word for word the same program
// for the same model (tests/test_js_host.py compares them). The closure is made traceable by amwg_rewrite.js (operators -> calls
// on the runtime object `__r` defined here); it is then run ONCE with symbolic parameters and proxied data:
//   * the returned value is split along its left spine of `+` into terms, in the order of the `log_post += ...` statements;
//   * runs of structurally identical terms that walk the data become plates (hand-written device loops for the recognised
//     bodies, a bytecode loop otherwise);
//   * keys the closure adds to `state` become derived quantities (tests/test_data.js:89);
//   * `if (m === 0)` on a BINARY parameter (tests/test_data.js:163-168) is recorded once per configuration of the binary
//     components; control flow on a real / int parameter cannot be recorded and throws (mcmc.where(c, a, b) is the device select).
(function (root, factory) {
  if (typeof define === "function" && define.amd) { define(["./amwg_rewrite"], factory); }
  else if (typeof module === "object" && module.exports) { module.exports = factory(typeof require === "function" ? require("./amwg_rewrite") : root.amwg_rewrite); }
  else { root.amwg_trace = factory(root.amwg_rewrite); }
}(this, function (rewriter) {
  "use strict";

  // ------------------------------------------------------------------------------------------------ opcodes (amwg.h)
  var OPS = ("END CONST COMP DATA DATA_I COMP_I ADD SUB MUL DIV NEG LOG EXP SQRT ABS POW LT LE GT GE EQ NE AND OR NOT SELECT " +
             "LGAMMA LFACTORIAL LCHOOSE LBETA LD_NORM LD_UNIF LD_BETA LD_BERN LD_POIS LD_CAUCHY LD_LAPLACE LD_GAMMA LD_INVGAMMA " +
             "LD_LNORM LD_PARETO LD_T LD_WEIBULL LD_LOGIS LD_EXP LD_BINOM LD_NBINOM LD_HYPER ACC PLATE STORE LOOP_BEGIN LOOP_END " +
             "NORM_K UNIF_K BETA_K ACC_RANGE PLATE_SS NORM_SS CACHED CAND").split(" ");
  var OP = {}, oi;
  for (oi = 0; oi < OPS.length; oi++) { OP[OPS[oi]] = oi; }
  var PLATE_GENERIC = 0, PLATE_NORM_IID = 1, PLATE_BERN_IID = 2, PLATE_NORM_GROUPED = 3;
  var MODE_STACK = 0, MODE_CONST = 1, MODE_COMP = 2, MODE_NONE = 3;
  var MAX_IMMEDIATE = 16383, STORE_FLAG = 131072, ACC_FLAG = 65536, MIN_PLATE = 8, MIN_STAT_POINTS = 64, MAX_VARIANT_COMPS = 4;
  var ARITY = {ADD: 2, SUB: 2, MUL: 2, DIV: 2, NEG: 1, LOG: 1, EXP: 1, SQRT: 1, ABS: 1, POW: 2, LT: 2, LE: 2, GT: 2, GE: 2, EQ: 2, NE: 2,
               AND: 2, OR: 2, NOT: 1, SELECT: 3, LGAMMA: 1, LFACTORIAL: 1, LCHOOSE: 2, LBETA: 2, LD_NORM: 3, LD_UNIF: 3, LD_BETA: 3,
               LD_BERN: 2, LD_POIS: 2, LD_CAUCHY: 3, LD_LAPLACE: 3, LD_GAMMA: 3, LD_INVGAMMA: 3, LD_LNORM: 3, LD_PARETO: 3, LD_T: 4,
               LD_WEIBULL: 3, LD_LOGIS: 3, LD_EXP: 2, LD_BINOM: 3, LD_NBINOM: 3, LD_HYPER: 4, NORM_K: 4, UNIF_K: 4, BETA_K: 4};

  function NeedsConcrete(message) { this.message = message; }

  // ------------------------------------------------------------------------------------------------ symbolic values
  function Sym(op, args, val) { this.op = op; this.args = args || []; this.val = val; }
  function is_sym(x) { return x instanceof Sym; }
  function is_num(x) { return typeof x === "number"; }
  function lift(x) {
    if (is_sym(x)) { return x; }
    if (x === true) { return new Sym("CONST", [], 1.0); }
    if (x === false) { return new Sym("CONST", [], 0.0); }
    if (is_num(x)) { return new Sym("CONST", [], x); }
    throw "log_post produced a value of type " + (typeof x) + " that is not a number";
  }
  function c_(v) { return new Sym("CONST", [], v); }
  function bin(op, a, b) { return new Sym(op, [lift(a), lift(b)]); }
  function un(op, a) { return new Sym(op, [lift(a)]); }
  function where(cond, a, b) {
    if (!is_sym(cond) && !is_sym(a) && !is_sym(b)) { return cond ? a : b; }
    return new Sym("SELECT", [lift(cond), lift(a), lift(b)]);
  }
  function num_key(v) { return v !== v ? "NaN" : (v === 0 && 1 / v < 0) ? "-0" : String(v); }

  // ------------------------------------------------------------------------------------------------ proxies
  function DataVec(t, col, shape, off) { this.t = t; this.col = col; this.shape = shape; this.off = off || 0; }
  DataVec.prototype.inner = function () { var n = 1, i; for (i = 1; i < this.shape.length; i++) { n *= this.shape[i]; } return n; };
  DataVec.prototype.get = function (key) {
    var i;
    if (key === "length") { return this.shape[0]; }
    if (is_sym(key)) { throw "indexing data by a parameter value is not supported on the device"; }
    i = Number(key);
    if (!(i === Math.floor(i)) || i < 0 || i >= this.shape[0]) { return undefined; }
    if (this.shape.length === 1) { return new Sym("DATA", [], [this.col, this.off + i]); }
    return new DataVec(this.t, this.col, this.shape.slice(1), this.off + i * this.inner());
  };
  function ParamVec(t, c0, shape) { this.t = t; this.c0 = c0; this.shape = shape; }
  ParamVec.prototype.get = function (key) {
    var i, inner = 1, k, v;
    if (key === "length") { return this.shape[0]; }
    for (k = 1; k < this.shape.length; k++) { inner *= this.shape[k]; }
    if (is_sym(key)) {
      if (key.op !== "DATA") { throw "a parameter array can only be indexed by numbers or by data values"; }
      v = this.t.columns[key.val[0]][key.val[1]];                 // concrete data value used as an index (mu[g[i]])
      if (v !== Math.floor(v)) { return new Sym("CONST", [], NaN); }
      i = v;
    } else { i = Number(key); }
    if (!(i === Math.floor(i)) || i < 0 || i >= this.shape[0]) { return new Sym("CONST", [], NaN); }
    if (this.shape.length === 1) { return this.t.comp(this.c0 + i); }
    return new ParamVec(this.t, this.c0 + i * inner, this.shape.slice(1));
  };
  function State() { this.keys = []; this.vals = {}; }
  State.prototype.get = function (key) {
    if (this.vals.hasOwnProperty(key)) { return this.vals[key]; }
    return this.base ? this.base[key] : undefined;                // stand-alone steppers: the rest of the caller's state object
  };
  State.prototype.set = function (key, v) { if (!this.vals.hasOwnProperty(key)) { this.keys.push(key); } this.vals[key] = v; return v; };

  function Tracer() { this.columns = []; this.n_plate_idx = 0; this.plate_sizes = {}; this.concrete = {}; }
  Tracer.prototype.comp = function (c) { return this.concrete.hasOwnProperty(c) ? this.concrete[c] : new Sym("COMP", [], c); };
  Tracer.prototype.add_column = function (arr) { this.columns.push(arr); return this.columns.length - 1; };
  function flat_numbers(a, out, shape, depth) {           // nested rectangular array of numbers -> flat; false when ragged / not numeric
    var i;
    if (Object.prototype.toString.call(a) === "[object Array]" || (typeof Float64Array !== "undefined" && a instanceof Float64Array)) {
      if (shape.length <= depth) { shape.push(a.length); } else if (shape[depth] !== a.length) { return false; }
      for (i = 0; i < a.length; i++) { if (!flat_numbers(a[i], out, shape, depth + 1)) { return false; } }
      return true;
    }
    if (typeof a === "boolean") { a = a ? 1 : 0; }
    if (!is_num(a) || shape.length !== depth) { return false; }
    out.push(a);
    return true;
  }
  Tracer.prototype.wrap_data = function (data) {
    var out, shape, k, o, i;
    if (data === null || data === undefined || typeof data === "string" || typeof data === "function") { return data; }
    if (typeof data === "boolean") { return data ? 1 : 0; }
    if (is_num(data)) { return data; }
    if (Object.prototype.toString.call(data) === "[object Array]" || (typeof Float64Array !== "undefined" && data instanceof Float64Array)) {
      out = []; shape = [];
      if (data.length === 0) { return []; }
      if (flat_numbers(data, out, shape, 0)) { return new DataVec(this, this.add_column(out), shape, 0); }
      o = [];
      for (i = 0; i < data.length; i++) { o.push(this.wrap_data(data[i])); }
      return o;
    }
    o = {};
    for (k in data) { if (data.hasOwnProperty(k)) { o[k] = this.wrap_data(data[k]); } }
    return o;
  };
  Tracer.prototype.make_state = function (names, params, offsets) {
    var st = new State(), i, p;
    for (i = 0; i < names.length; i++) {
      p = params[names[i]];
      if (p.dim.length === 1 && p.dim[0] === 1) { st.set(names[i], this.comp(offsets[names[i]])); }
      else { st.set(names[i], new ParamVec(this, offsets[names[i]], p.dim.slice(0))); }
    }
    return st;
  };

  // ------------------------------------------------------------------------------------------------ the runtime `__r`
  function make_runtime(native_ld, active) {
    var R = {};
    function arith(op, f) {
      return function (a, b) { if (is_sym(a) || is_sym(b)) { return bin(op, a, b); } return f(a, b); };
    }
    R.add = arith("ADD", function (a, b) { return a + b; });
    R.sub = arith("SUB", function (a, b) { return a - b; });
    R.mul = arith("MUL", function (a, b) { return a * b; });
    R.div = arith("DIV", function (a, b) { return a / b; });
    R.mod = function (a, b) { if (is_sym(a) || is_sym(b)) { throw "the % operator on a parameter value is not supported on the device"; } return a % b; };
    R.lt = arith("LT", function (a, b) { return a < b; });
    R.le = arith("LE", function (a, b) { return a <= b; });
    R.gt = arith("GT", function (a, b) { return a > b; });
    R.ge = arith("GE", function (a, b) { return a >= b; });
    R.eq = arith("EQ", function (a, b) { return a === b; });
    R.ne = arith("NE", function (a, b) { return a !== b; });
    R.neg = function (a) { return is_sym(a) ? un("NEG", a) : -a; };
    R.pos = function (a) { return is_sym(a) ? a : +a; };
    R.t = function (v) {
      if (is_sym(v)) { throw new NeedsConcrete("log_post branches on a parameter value, which cannot be traced for the device; use mcmc.where(cond, a, b)"); }
      return !!v;
    };
    R.concrete = function (v) {
      if (is_sym(v)) { throw new NeedsConcrete("log_post uses a parameter value where JavaScript needs a concrete one"); }
      return v;
    };
    R.and = function (a, fb) { if (is_sym(a)) { return new Sym("AND", [a, lift(fb())]); } return a ? fb() : a; };
    R.or = function (a, fb) { if (is_sym(a)) { return new Sym("OR", [a, lift(fb())]); } return a ? a : fb(); };
    R.get = function (obj, key) {
      if (obj instanceof DataVec || obj instanceof ParamVec || obj instanceof State) { return obj.get(key); }
      if (is_sym(obj)) { throw "a parameter value has no property " + key; }
      if (is_sym(key)) { throw new NeedsConcrete("log_post uses a parameter value as an index; only binary parameters can be used that way"); }
      return obj[key];
    };
    R.set = function (obj, key, v) {
      if (obj instanceof State) { return obj.set(key, v); }
      if (obj instanceof DataVec || obj instanceof ParamVec) { throw "log_post may not write into data or parameter arrays"; }
      obj[key] = v;
      return v;
    };
    R.del = function (obj, key) { if (obj instanceof State) { throw "deleting state keys is not supported"; } return delete obj[key]; };
    R.call = function (obj, name, args) {
      var f = R.get(obj, name);
      if (typeof f !== "function") { throw "TypeError: " + name + " is not a function"; }
      return f.apply(obj, args);
    };
    R.keys = function (obj) {
      var o = {}, i;
      if (obj instanceof State) { for (i = 0; i < obj.keys.length; i++) { o[obj.keys[i]] = 1; } return o; }
      if (obj instanceof DataVec || obj instanceof ParamVec) { for (i = 0; i < obj.shape[0]; i++) { o[i] = 1; } return o; }
      return obj;
    };
    function m1(op, f) { return function (x) { return is_sym(x) ? un(op, x) : f(x); }; }
    R.Math = {
      PI: Math.PI, E: Math.E, LN2: Math.LN2, LN10: Math.LN10, SQRT2: Math.SQRT2,
      log: m1("LOG", Math.log), exp: m1("EXP", Math.exp), sqrt: m1("SQRT", Math.sqrt), abs: m1("ABS", Math.abs),
      pow: function (x, y) { return (is_sym(x) || is_sym(y)) ? bin("POW", x, y) : Math.pow(x, y); },
      max: function (a, b) {
        if (!is_sym(a) && !is_sym(b)) { return Math.max(a, b); }
        a = lift(a); b = lift(b);
        return where(new Sym("OR", [bin("NE", a, a), bin("NE", b, b)]), NaN, where(bin("GT", a, b), a, b));
      },
      min: function (a, b) {
        if (!is_sym(a) && !is_sym(b)) { return Math.min(a, b); }
        a = lift(a); b = lift(b);
        return where(new Sym("OR", [bin("NE", a, a), bin("NE", b, b)]), NaN, where(bin("LT", a, b), a, b));
      },
      floor: function (x) { return Math.floor(R.concrete(x)); }, ceil: function (x) { return Math.ceil(R.concrete(x)); },
      round: function (x) { return Math.round(R.concrete(x)); }, random: function () { throw "log_post must be a pure function of the state: Math.random() inside it cannot be recorded"; }
    };
    // ld.* inside log_post: every call records one node (constants are folded on the device), as in tracer.py / distributions.py
    function ldop(op, arity, name) {
      return function () {
        var args = [], i;
        if (arguments.length !== arity) { throw "ld." + name + " takes " + arity + " arguments"; }
        for (i = 0; i < arity; i++) { args.push(lift(arguments[i])); }
        return new Sym(op, args);
      };
    }
    R.ld = {
      lgamma: ldop("LGAMMA", 1, "lgamma"), lfactorial: ldop("LFACTORIAL", 1, "lfactorial"), lchoose: ldop("LCHOOSE", 2, "lchoose"),
      lbeta: ldop("LBETA", 2, "lbeta"), beta: ldop("LD_BETA", 3, "beta"), cauchy: ldop("LD_CAUCHY", 3, "cauchy"), norm: ldop("LD_NORM", 3, "norm"),
      laplace: ldop("LD_LAPLACE", 3, "laplace"), dexp: ldop("LD_LAPLACE", 3, "dexp"), gamma: ldop("LD_GAMMA", 3, "gamma"),
      invgamma: ldop("LD_INVGAMMA", 3, "invgamma"), lnorm: ldop("LD_LNORM", 3, "lnorm"), pareto: ldop("LD_PARETO", 3, "pareto"),
      t: ldop("LD_T", 4, "t"), weibull: ldop("LD_WEIBULL", 3, "weibull"), logis: ldop("LD_LOGIS", 3, "logis"), exp: ldop("LD_EXP", 2, "exp"),
      unif: ldop("LD_UNIF", 3, "unif"), bern: ldop("LD_BERN", 2, "bern"), binom: ldop("LD_BINOM", 3, "binom"), nbinom: ldop("LD_NBINOM", 3, "nbinom"),
      hyper: ldop("LD_HYPER", 4, "hyper"), pois: ldop("LD_POIS", 2, "pois"),
      bivarnorm: function (x, mean, sd, corr) {                   // distributions.js:125-133, composed from primitives in the JS order
        var g = R.get, M = R.Math, x0 = lift(g(x, 0)), x1 = lift(g(x, 1)), m0 = lift(g(mean, 0)), m1_ = lift(g(mean, 1)), s0 = lift(g(sd, 0)), s1 = lift(g(sd, 1)), r = lift(corr);
        var z = bin("SUB", bin("ADD", bin("DIV", M.pow(bin("SUB", x0, m0), 2), M.pow(s0, 2)), bin("DIV", M.pow(bin("SUB", x1, m1_), 2), M.pow(s1, 2))),
                    bin("DIV", bin("MUL", bin("MUL", bin("MUL", 2, r), bin("SUB", x0, m0)), bin("SUB", x1, m1_)), bin("MUL", s0, s1)));
        var nf = un("NEG", bin("ADD", bin("ADD", bin("ADD", bin("ADD", M.log(c_(2)), M.log(c_(Math.PI))), M.log(s0)), M.log(s1)), bin("MUL", 0.5, M.log(bin("SUB", 1, M.pow(r, 2))))));
        return bin("SUB", nf, bin("DIV", z, bin("MUL", 2, bin("SUB", 1, M.pow(r, 2)))));
      },
      dirichlet: function (x, alpha) {                            // distributions.js:203-214
        var n = R.get(alpha, "length"), i, a, sum_alpha = c_(0), sum_lg = c_(0), s = c_(0);
        for (i = 0; i < n; i++) {
          a = lift(R.get(alpha, i));
          sum_alpha = bin("ADD", sum_alpha, a);
          sum_lg = bin("ADD", sum_lg, new Sym("LGAMMA", [a]));
          s = bin("ADD", s, bin("MUL", bin("SUB", a, 1), R.Math.log(lift(R.get(x, i)))));
        }
        return bin("ADD", bin("SUB", new Sym("LGAMMA", [sum_alpha]), sum_lg), s);
      },
      cat: function (x, probs) {                                  // distributions.js:232-238: probs[x - 1] as a SELECT chain
        var xs = lift(x), n = R.get(probs, "length"), k, picked = c_(NaN);
        for (k = n; k > 0; k--) { picked = where(bin("EQ", xs, k), R.Math.log(lift(R.get(probs, k - 1))), picked); }
        return where(new Sym("OR", [bin("LT", xs, 1), bin("GT", xs, n)]), -Infinity, picked);
      }
    };
    R.where = where;
    return R;
  }

  // ------------------------------------------------------------------------------------------------ lowering helpers
  function is_const(n) {
    var stack = [n], m, i;
    while (stack.length) {
      m = stack.pop();
      if (m.op === "COMP" || m.op === "DATA_I" || m.op === "COMP_I") { return false; }
      for (i = 0; i < m.args.length; i++) { stack.push(m.args[i]); }
    }
    return true;
  }
  function const_value(n) { return n.op === "CONST" ? n.val : null; }
  var ML = function (x) { return un("LOG", x); };
  function expand_ld(op, a) {
    var x, mean, sd, k1, k2, mn, mx, k, s1, s2, body, p, lam, rate, loc, scale;
    if (op === "LD_NORM") {                                       // distributions.js:119-121
      x = a[0]; mean = a[1]; sd = a[2];
      k1 = bin("SUB", bin("MUL", c_(-0.5), ML(bin("MUL", c_(2), c_(Math.PI)))), ML(sd));
      k2 = bin("MUL", bin("MUL", c_(2), sd), sd);
      if (is_const(sd)) { return new Sym("NORM_K", [x, mean, k1, k2]); }
      return bin("SUB", k1, bin("DIV", bin("POW", bin("SUB", x, mean), 2), k2));
    }
    if (op === "LD_UNIF") {                                       // :221-223
      x = a[0]; mn = a[1]; mx = a[2];
      k = ML(bin("DIV", c_(1), bin("SUB", mx, mn)));
      if (is_const(mn) && is_const(mx)) { return new Sym("UNIF_K", [x, mn, mx, k]); }
      return where(new Sym("OR", [bin("LT", x, mn), bin("GT", x, mx)]), -Infinity, k);
    }
    if (op === "LD_BETA") {                                       // :104-113
      x = a[0]; s1 = a[1]; s2 = a[2];
      if (const_value(s1) !== null && const_value(s2) !== null) {
        if (const_value(s1) === 1 && const_value(s2) === 1) { return where(new Sym("OR", [bin("GT", x, 1), bin("LT", x, 0)]), -Infinity, 0.0); }
        return new Sym("BETA_K", [x, bin("SUB", s1, 1), bin("SUB", s2, 1), new Sym("LBETA", [s1, s2])]);
      }
      body = bin("SUB", bin("ADD", bin("MUL", bin("SUB", s1, 1), ML(x)), bin("MUL", bin("SUB", s2, 1), ML(bin("SUB", c_(1), x)))), new Sym("LBETA", [s1, s2]));
      return where(new Sym("OR", [bin("GT", x, 1), bin("LT", x, 0)]), -Infinity, where(new Sym("AND", [bin("EQ", s1, 1), bin("EQ", s2, 1)]), 0.0, body));
    }
    if (op === "LD_BERN") {                                       // :228-230
      x = a[0]; p = a[1];
      return where(new Sym("NOT", [new Sym("OR", [bin("EQ", x, 0), bin("EQ", x, 1)])]), -Infinity,
                   ML(bin("ADD", bin("MUL", x, p), bin("MUL", bin("SUB", c_(1), x), bin("SUB", c_(1), p)))));
    }
    if (op === "LD_POIS") {                                       // :282-284
      x = a[0]; lam = a[1];
      return where(bin("LT", x, 0), -Infinity, bin("SUB", bin("SUB", bin("MUL", ML(lam), x), lam), new Sym("LFACTORIAL", [x])));
    }
    if (op === "LD_EXP") { x = a[0]; rate = a[1]; return where(bin("LT", x, 0), -Infinity, bin("SUB", ML(rate), bin("MUL", rate, x))); }   // :217-219
    if (op === "LD_LAPLACE") {                                    // :136-138
      x = a[0]; loc = a[1]; scale = a[2];
      return bin("SUB", bin("DIV", un("NEG", un("ABS", bin("SUB", x, loc))), scale), ML(bin("MUL", c_(2), scale)));
    }
    if (op === "LD_CAUCHY") {                                     // :115-117
      x = a[0]; loc = a[1]; scale = a[2];
      return bin("SUB", bin("SUB", ML(scale), ML(bin("ADD", bin("POW", bin("SUB", x, loc), 2), bin("POW", scale, 2)))), ML(c_(Math.PI)));
    }
    return null;
  }
  function expand(node) {
    var args = [], i, e;
    if (!node.args.length) { return node; }
    for (i = 0; i < node.args.length; i++) { args.push(expand(node.args[i])); }
    if (node.op.substring(0, 3) === "LD_") { e = expand_ld(node.op, args); if (e !== null) { return e; } }
    return new Sym(node.op, args, node.val);
  }
  function spine_terms(expr) {
    var terms = [], node = expr;
    while (node.op === "ADD") { terms.push(node.args[1]); node = node.args[0]; }
    if (!(node.op === "CONST" && node.val === 0 && terms.length)) { terms.push(node); }
    terms.reverse();
    return terms;
  }
  function signature(node, slots, loose) {
    var i, out;
    if (node.op === "CONST") { return "K" + num_key(node.val); }
    if (node.op === "COMP") { if (loose) { slots.push(["C", -1, node.val]); return "C?"; } return "C" + node.val; }
    if (node.op === "DATA") { slots.push(["D", node.val[0], node.val[1]]); return "D" + node.val[0]; }
    if (node.op === "DATA_I" || node.op === "COMP_I") { return node.op + "<" + node.val.join(",") + ">"; }
    out = node.op + "(";
    for (i = 0; i < node.args.length; i++) { out += signature(node.args[i], slots, loose) + ";"; }
    return out + ")";
  }
  function has_plate_ref(node) {
    var stack = [node], n, i;
    while (stack.length) {
      n = stack.pop();
      if (n.op === "DATA_I") { return n.val[3]; }
      if (n.op === "COMP_I") { return n.val[4]; }
      for (i = 0; i < n.args.length; i++) { stack.push(n.args[i]); }
    }
    return null;
  }
  function index_free(node) {
    var stack = [node], n, i;
    while (stack.length) {
      n = stack.pop();
      if (n.op === "DATA_I" || n.op === "COMP_I") { return false; }
      for (i = 0; i < n.args.length; i++) { stack.push(n.args[i]); }
    }
    return true;
  }
  function set_keys(s) { var k, out = []; for (k in s) { if (s.hasOwnProperty(k)) { out.push(Number(k)); } } return out; }
  function set_union(a, b) { var k, o = {}; for (k in a) { if (a.hasOwnProperty(k)) { o[k] = 1; } } for (k in b) { if (b.hasOwnProperty(k)) { o[k] = 1; } } return o; }
  function set_size(s) { return set_keys(s).length; }

  // ------------------------------------------------------------------------------------------------ program
  function Program() {
    this.code = []; this.consts = []; this.const_index = {}; this.columns = []; this.plates = [];
    this.logpost_prog = 0; this.derived_prog = -1; this.derived_names = []; this.store_sites = [];
    this.n_terms = 0; this.comp_prog = []; this.touch_off = []; this.touch_terms = [];
    this.stat_prog = -1; this.n_sum_terms = 0; this.block_params = []; this.term_block_comp = [];
    this.variant_comps = []; this.variant_logpost = []; this.variant_derived = [];
    this.fold_prog = []; this.fold_dst = []; this.summary = [];
  }
  Program.prototype.const_ = function (v) {
    var key = num_key(v);
    if (!this.const_index.hasOwnProperty(key)) { this.const_index[key] = this.consts.length; this.consts.push(v); }
    return this.const_index[key];
  };
  Program.prototype.fold_slot = function () { this.consts.push(NaN); return this.consts.length - 1; };
  Program.prototype.emit = function (op, operand, extra, modes, acc, store) {
    var m = [MODE_NONE, MODE_NONE, MODE_NONE, MODE_NONE], i, word;
    operand = operand || 0; extra = extra || []; modes = modes || [];
    if (!(operand >= 0 && operand <= MAX_IMMEDIATE)) { throw "log_post is too large for the device program format (immediate > 16383)"; }
    for (i = 0; i < modes.length; i++) { m[i] = modes[i]; }
    // bits 0-7 opcode, 8-15 operand modes, 16 ACC, 17 STORE, 18-31 immediate (as a signed 32-bit word, like the C side)
    word = OP[op] + m[0] * 256 + m[1] * 1024 + m[2] * 4096 + m[3] * 16384 + (acc ? ACC_FLAG : 0) + (store !== null && store !== undefined ? STORE_FLAG : 0) + operand * 262144;
    if (word >= 2147483648) { word -= 4294967296; }
    this.code.push(word);
    for (i = 0; i < extra.length; i++) { this.code.push(extra[i]); }
    if (store !== null && store !== undefined) {
      this.store_sites.push([this.code.length - 1 - extra.length, this.code.length]);
      this.code.push(store);
    }
  };
  function word_or(w, flag) { var u = w < 0 ? w + 4294967296 : w; if (Math.floor(u / flag) % 2 === 0) { u += flag; } return u >= 2147483648 ? u - 4294967296 : u; }
  function word_clear(w, flag) { var u = w < 0 ? w + 4294967296 : w; if (Math.floor(u / flag) % 2 === 1) { u -= flag; } return u >= 2147483648 ? u - 4294967296 : u; }

  function Lowering(tracer, n_comp, faithful, param_ranges) {
    this.t = tracer; this.n_comp = n_comp; this.faithful = !!faithful; this.param_ranges = param_ranges || [];
    this.prog = new Program(); this.prog.columns = tracer.columns;
    this.fold_memo = {}; this.fold_trees = []; this.terms = []; this.abs_words = [];
    this.record_terms = false; this.stat_mode = false; this.stat_lowering = true;
  }
  Lowering.prototype.key = function (n) {
    var i, out;
    if (n.op === "CONST") { return "K" + num_key(n.val); }
    if (n.op === "COMP" || n.op === "FOLD") { return n.op + n.val; }
    if (n.op === "DATA" || n.op === "DATA_I" || n.op === "COMP_I") { return n.op + "<" + n.val.join(",") + ">"; }
    out = n.op + "(";
    for (i = 0; i < n.args.length; i++) { out += this.key(n.args[i]) + ";"; }
    return out + ")";
  };
  Lowering.prototype.to_fold = function (tree) {
    var key = this.key(tree), k;
    if (!this.fold_memo.hasOwnProperty(key)) { k = this.prog.fold_slot(); this.fold_memo[key] = k; this.fold_trees.push([k, tree]); }
    return new Sym("FOLD", [], this.fold_memo[key]);
  };
  Lowering.prototype.fold = function (node) {
    var self = this;
    function rec(n) {
      var parts = [], i, all = true, args = [];
      if (n.op === "CONST" || n.op === "DATA" || n.op === "FOLD") { return [n, true]; }
      if (n.op === "COMP" || n.op === "DATA_I" || n.op === "COMP_I") { return [n, false]; }
      for (i = 0; i < n.args.length; i++) { parts.push(rec(n.args[i])); all = all && parts[i][1]; }
      if (all) { for (i = 0; i < parts.length; i++) { args.push(parts[i][0]); } return [new Sym(n.op, args, n.val), true]; }
      for (i = 0; i < parts.length; i++) { args.push((parts[i][1] && parts[i][0].args.length) ? self.to_fold(parts[i][0]) : parts[i][0]); }
      return [new Sym(n.op, args, n.val), false];
    }
    var r = rec(node);
    return (r[1] && r[0].args.length) ? this.to_fold(r[0]) : r[0];
  };
  Lowering.prototype.inline = function (n) {
    if (n.op === "CONST") { return [MODE_CONST, this.prog.const_(n.val)]; }
    if (n.op === "FOLD") { return [MODE_CONST, n.val]; }
    if (n.op === "COMP") { return [MODE_COMP, n.val]; }
    return null;
  };
  Lowering.prototype.operands = function (args) {
    var modes = [], words = [], i, il;
    for (i = 0; i < args.length; i++) {
      il = this.inline(args[i]);
      if (il === null) { this.emit_(args[i], false, null); modes.push(MODE_STACK); } else { modes.push(il[0]); words.push(il[1]); }
    }
    words.reverse();
    return [modes, words];
  };
  Lowering.prototype.emit_ = function (n, acc, store) {
    var p = this.prog, il, mw;
    if (n.op === "CONST" || n.op === "FOLD" || n.op === "COMP") {
      il = this.inline(n);
      p.emit(il[0] === MODE_CONST ? "CONST" : "COMP", il[1], [], [], acc, store);
    } else if (n.op === "DATA") { p.emit("DATA", n.val[0], [n.val[1]], [], acc, store); }
    else if (n.op === "DATA_I") { p.emit("DATA_I", n.val[0], [n.val[1], n.val[2]], [], acc, store); }
    else if (n.op === "COMP_I") { p.emit("COMP_I", n.val[0], [n.val[1], n.val[2], n.val[3]], [], acc, store); }
    else {
      if (!ARITY.hasOwnProperty(n.op)) { throw "cannot lower operation " + n.op; }
      mw = this.operands(n.args);
      p.emit(n.op, 0, mw[1], mw[0], acc, store);
    }
  };
  Lowering.prototype.prepared = function (node) { return this.fold(expand(node)); };
  Lowering.prototype.emit_expr = function (node, prepare, acc, store) {
    if (prepare !== false) { node = this.prepared(node); }
    this.emit_(node, !!acc, store === undefined ? null : store);
  };
  Lowering.prototype.deps = function (node) {
    var out = {}, stack = [node], n, i, v, cnt, col;
    while (stack.length) {
      n = stack.pop();
      if (n.op === "COMP") { out[n.val] = 1; }
      else if (n.op === "COMP_I") {
        v = n.val; cnt = this.t.plate_sizes[v[4]] || 0; col = this.t.columns[v[0]];
        if (cnt) { for (i = 0; i < cnt; i++) { out[v[3] + col[v[1] + v[2] * i]] = 1; } }
        else { for (i = 0; i < col.length; i++) { out[v[3] + col[i]] = 1; } }
      }
      for (i = 0; i < n.args.length; i++) { stack.push(n.args[i]); }
    }
    return out;
  };

  // ---- plates
  Lowering.prototype.grouped = function (mean, n) {
    var v = mean.val, col = v[0], off = v[1], stride = v[2], base = v[3], g, i, lo, hi, J, start = [], j, pos;
    if (stride !== 1) { return null; }
    g = this.t.columns[col];
    for (i = 0; i < n; i++) { if (g[off + i] !== Math.floor(g[off + i]) || (i > 0 && g[off + i] < g[off + i - 1])) { return null; } }
    lo = g[off]; hi = g[off + n - 1]; J = hi - lo + 1;
    pos = 0;
    for (j = lo; j <= hi + 1; j++) { while (pos < n && g[off + pos] < j) { pos++; } start.push(pos); }
    return [base + lo, J, this.t.add_column(start)];
  };
  Lowering.prototype.emit_plate = function (body, n) {
    var p = this.prog, pl = {kind: PLATE_GENERIC, n: n, col: [-1, -1, -1, -1], iparam: [0, 0, 0, 0]}, q = p.plates.length, operands = [];
    var x, mean, sd, grp, start, mean_p, sd_p, k, i, mw, ss_word, split, m_sd, prepared, value_plate, tid, deps, fix, body_start;
    function data_i(node) { return node.op === "DATA_I" && node.val[2] === 1; }
    if (this.faithful && (body.op === "LD_NORM" || body.op === "LD_POIS")) {
      x = null;                                                   // bytecode loop below: bit-faithful to the reference's arithmetic
    } else if (body.op === "LD_NORM" && data_i(body.args[0]) && index_free(body.args[2])) {
      x = body.args[0]; mean = body.args[1]; sd = body.args[2];
      if (index_free(mean)) {
        pl.kind = PLATE_NORM_IID; pl.col[0] = x.val[0]; pl.iparam[2] = x.val[1];
        operands = [mean, sd];
        p.summary.push("plate NORM_IID n=" + n);
      } else if (mean.op === "COMP_I") {
        grp = this.grouped(mean, n);
        if (grp !== null) {
          pl.kind = PLATE_NORM_GROUPED; pl.col[0] = x.val[0]; pl.col[1] = grp[2]; pl.iparam[2] = x.val[1]; pl.iparam[0] = grp[0]; pl.iparam[1] = grp[1];
          operands = [sd];
          p.summary.push("plate NORM_GROUPED n=" + n + " groups=" + grp[1]);
        }
      }
    } else if (body.op === "LD_BERN" && data_i(body.args[0]) && index_free(body.args[1])) {
      pl.kind = PLATE_BERN_IID; pl.col[0] = body.args[0].val[0]; pl.iparam[2] = body.args[0].val[1];
      operands = [body.args[1]];
      p.summary.push("plate BERN_IID n=" + n);
    }
    p.plates.push(pl);
    start = p.code.length;
    if (pl.kind === PLATE_NORM_IID && this.stat_mode) {
      mean_p = this.prepared(operands[0]); sd_p = this.prepared(operands[1]);
      k = 0;
      for (i = 0; i < this.terms.length; i++) { if (this.terms[i].stat) { k++; } }
      mw = this.operands([mean_p]);
      ss_word = p.code.length;
      p.emit("PLATE_SS", q, mw[1].concat([-1 - k]), mw[0], false, null);
      split = p.code.length;
      m_sd = this.operands([sd_p]);
      p.emit("NORM_SS", q, m_sd[1], [MODE_STACK].concat(m_sd[0]), true, this.terms.length);
      this.terms.push({start: start, end: p.code.length, kind: "value", deps: set_union(this.deps(mean_p), this.deps(sd_p)), cost: 3 * n,
                       stat: {k: k, ss_word: ss_word, slot_word: split - 1, split: split, mean_deps: this.deps(mean_p), n: n}});
      return;
    }
    if (pl.kind !== PLATE_GENERIC) {
      prepared = [];
      for (i = 0; i < operands.length; i++) { prepared.push(this.prepared(operands[i])); }
      mw = this.operands(prepared);
      value_plate = pl.kind !== PLATE_BERN_IID;
      tid = (this.record_terms && value_plate) ? this.terms.length : null;
      p.emit("PLATE", q, mw[1], mw[0], false, tid);
      if (this.record_terms) {
        deps = {};
        for (i = 0; i < prepared.length; i++) { deps = set_union(deps, this.deps(prepared[i])); }
        if (pl.kind === PLATE_NORM_GROUPED) { for (i = 0; i < pl.iparam[1]; i++) { deps[pl.iparam[0] + i] = 1; } }
        this.terms.push({start: start, end: p.code.length, kind: value_plate ? "value" : "inorder", deps: deps, cost: 3 * n, plate_kind: pl.kind,
                         mean_deps: pl.kind === PLATE_NORM_IID ? this.deps(prepared[0]) : null, n: n});
      }
      return;
    }
    p.emit("LOOP_BEGIN", q, [0], [], false, null);
    fix = p.code.length - 1;
    this.abs_words.push(fix);
    body_start = p.code.length;
    this.emit_expr(body);
    p.emit("LOOP_END", 0, [body_start], [], false, null);
    this.abs_words.push(p.code.length - 1);
    p.code[fix] = p.code.length;
    if (this.record_terms) {
      this.terms.push({start: start, end: p.code.length, kind: "inorder", deps: {}, cost: 12 * n * Math.max(1, p.code.length - body_start), plate_kind: PLATE_GENERIC});
    }
    p.summary.push("plate GENERIC n=" + n + " body=" + body.op);
  };
  Lowering.prototype.rebuild = function (node, plan, loose) {
    var args = [], i, item;
    if (node.op === "DATA" || (node.op === "COMP" && loose)) { item = plan.shift(); return new Sym(item[0], [], item[1]); }
    if (!node.args.length) { return node; }
    for (i = 0; i < node.args.length; i++) { args.push(this.rebuild(node.args[i], plan, loose)); }
    return new Sym(node.op, args, node.val);
  };
  Lowering.prototype.find_run = function (terms, i0) {
    var slots0 = [], sig0 = signature(terms[i0], slots0, false), loose = false, s1 = [], seqs = [], j, sl, k, length, plan = [], arr, stride, brk, pid, final_plan = [], base, col, same, ok;
    if (!slots0.length || i0 + 1 >= terms.length) { return null; }
    if (signature(terms[i0 + 1], s1, false) !== sig0) {
      slots0 = []; sig0 = signature(terms[i0], slots0, true); s1 = [];
      if (signature(terms[i0 + 1], s1, true) !== sig0) { return null; }
      loose = true;
    }
    for (k = 0; k < slots0.length; k++) { seqs.push([slots0[k][2]]); }
    j = i0 + 1;
    while (j < terms.length) {
      sl = [];
      if (signature(terms[j], sl, loose) !== sig0 || sl.length !== slots0.length) { break; }
      ok = true;
      for (k = 0; k < sl.length; k++) { if (sl[k][0] !== slots0[k][0] || sl[k][1] !== slots0[k][1]) { ok = false; } }
      if (!ok) { break; }
      for (k = 0; k < sl.length; k++) { seqs[k].push(sl[k][2]); }
      j++;
    }
    length = j - i0;
    if (length < MIN_PLATE) { return null; }
    for (k = 0; k < slots0.length; k++) {
      arr = seqs[k];
      if (slots0[k][0] === "D") {
        stride = arr[1] - arr[0];
        brk = -1;
        for (j = 1; j < arr.length; j++) { if (arr[j] - arr[j - 1] !== stride) { brk = j; break; } }
        if (brk >= 0) { length = Math.min(length, brk); }
        plan.push(["D", slots0[k][1], arr[0], stride]);
      } else { plan.push(["C", arr]); }
    }
    if (length < MIN_PLATE) { return null; }
    pid = this.t.n_plate_idx++;
    this.t.plate_sizes[pid] = length;
    for (k = 0; k < plan.length; k++) {
      if (plan[k][0] === "D") { final_plan.push(["DATA_I", [plan[k][1], plan[k][2], plan[k][3], pid]]); }
      else {
        arr = plan[k][1].slice(0, length);
        same = true; base = arr[0];
        for (j = 0; j < arr.length; j++) { if (arr[j] !== arr[0]) { same = false; } if (arr[j] < base) { base = arr[j]; } }
        if (same) { final_plan.push(["COMP", arr[0]]); }
        else {
          col = [];
          for (j = 0; j < arr.length; j++) { col.push(arr[j] - base); }
          final_plan.push(["COMP_I", [this.t.add_column(col), 0, 1, base, pid]]);
        }
      }
    }
    return [this.rebuild(terms[i0], final_plan, loose), length];
  };

  // ---- programs
  Lowering.prototype.add_logpost = function (result, derived_names, derived) {
    var p = this.prog, lp_off = p.code.length, terms = spine_terms(result), i = 0, tm, pid, run, node, start, tid, der_off = -1, d, same;
    while (i < terms.length) {
      tm = terms[i];
      pid = has_plate_ref(tm);
      if (pid !== null) { this.emit_plate(tm, this.t.plate_sizes[pid]); i++; continue; }
      run = this.find_run(terms, i);
      if (run !== null) { this.emit_plate(run[0], run[1]); i += run[1]; continue; }
      node = this.prepared(tm);
      start = p.code.length;
      tid = this.record_terms ? this.terms.length : null;
      this.emit_expr(node, false, true, tid);
      if (this.record_terms) { this.terms.push({start: start, end: p.code.length, kind: "value", deps: this.deps(node), cost: 10 * (p.code.length - start)}); }
      p.summary.push("term " + tm.op);
      i++;
    }
    p.emit("END");
    if (derived_names.length) {
      der_off = p.code.length;
      if (p.derived_names.length) {
        same = p.derived_names.length === derived_names.length;
        for (d = 0; same && d < derived_names.length; d++) { same = p.derived_names[d] === derived_names[d]; }
        if (!same) { throw "log_post adds different derived quantities for different values of the binary parameters"; }
      }
      p.derived_names = derived_names.slice(0);
      for (d = 0; d < derived_names.length; d++) { this.emit_expr(lift(derived[derived_names[d]])); p.emit("STORE", d); }
      p.emit("END");
    }
    return [lp_off, der_off];
  };
  Lowering.prototype.finish = function () {
    var p = this.prog, i;
    for (i = 0; i < this.fold_trees.length; i++) {
      p.fold_prog.push(p.code.length);
      p.fold_dst.push(this.fold_trees[i][0]);
      this.emit_expr(this.fold_trees[i][1], false);
      p.emit("END");
    }
    return p;
  };
  Lowering.prototype.per_component_cost = function () {
    var out = [], c, t, tr, a, b;
    for (c = 0; c < this.n_comp; c++) {
      a = 0; b = 0;
      for (t = 0; t < this.terms.length; t++) {
        tr = this.terms[t];
        if (tr.kind === "inorder" || tr.deps.hasOwnProperty(c)) { a += tr.cost; }
        if (tr.kind === "value" && !tr.deps.hasOwnProperty(c)) { b++; }
      }
      out.push(a + 3 * b + 40);
    }
    return out;
  };
  Lowering.prototype.cache_worthwhile = function () {
    var any = false, full = 0, t, sum = 0, pc;
    for (t = 0; t < this.terms.length; t++) { any = any || this.terms[t].kind === "value"; full += this.terms[t].cost; }
    if (!any || this.terms.length > MAX_IMMEDIATE) { return false; }
    pc = this.per_component_cost();
    for (t = 0; t < pc.length; t++) { sum += pc[t]; }
    return sum <= 0.6 * full * this.n_comp;
  };
  Lowering.prototype.emit_component_programs = function () {
    var p = this.prog, terms = this.terms, full = 0, pc = this.per_component_cost(), sum = 0, c, t, tr, run_start, run_len, delta, frag, k, pos, all_value, pidx, r, comps, row, hit, hk, i;
    for (t = 0; t < terms.length; t++) { full += terms[t].cost; }
    for (t = 0; t < pc.length; t++) { sum += pc[t]; }
    var abs_words = this.abs_words.slice(0);
    abs_words.sort(function (a, b) { return a - b; });
    p.n_terms = terms.length;
    function flush() { if (run_len) { p.emit("ACC_RANGE", run_start, [run_len]); } run_start = null; run_len = 0; }
    for (c = 0; c < this.n_comp; c++) {
      p.comp_prog.push(p.code.length);
      p.touch_off.push(p.touch_terms.length);
      run_start = null; run_len = 0;
      for (t = 0; t < terms.length; t++) {
        tr = terms[t];
        if (tr.kind === "value" && !tr.deps.hasOwnProperty(c)) {
          if (run_len && run_start + run_len === t) { run_len++; } else { flush(); run_start = t; run_len = 1; }
          continue;
        }
        flush();
        delta = p.code.length - tr.start;
        frag = p.code.slice(tr.start, tr.end);
        for (k = 0; k < abs_words.length; k++) { pos = abs_words[k]; if (tr.start <= pos && pos < tr.end) { frag[pos - tr.start] += delta; } }
        for (k = 0; k < frag.length; k++) { p.code.push(frag[k]); }
        if (tr.kind === "value") { p.touch_terms.push(t); }
      }
      flush();
      p.emit("END");
    }
    p.touch_off.push(p.touch_terms.length);
    p.summary.push("dependency-aware evaluation: " + terms.length + " terms, cost " + (sum / (full * this.n_comp)).toFixed(2) + " of the full program");
    all_value = true;
    for (t = 0; t < terms.length; t++) { all_value = all_value && terms[t].kind === "value"; }
    if (all_value) {
      for (pidx = 0; pidx < this.param_ranges.length; pidx++) {
        r = this.param_ranges[pidx];
        if (r[1] <= 1 || r[2] === "binary" || p.block_params.length >= 4) { continue; }
        row = [];
        for (t = 0; t < terms.length && row !== null; t++) {
          hit = [];
          hk = set_keys(terms[t].deps);
          for (i = 0; i < hk.length; i++) { if (hk[i] >= r[0] && hk[i] < r[0] + r[1]) { hit.push(hk[i]); } }
          if (hit.length > 1) { row = null; } else { row.push(hit.length ? hit[0] : -1); }
        }
        if (row !== null) {
          p.block_params.push(pidx);
          for (i = 0; i < row.length; i++) { p.term_block_comp.push(row[i]); }
          p.summary.push("block steps for parameter #" + pidx + ": " + r[1] + " components with one evaluation");
        }
      }
    }
  };
  Lowering.prototype.strip_stores = function (der_off) {
    var p = this.prog, removed = [], i, self = this;
    for (i = 0; i < p.store_sites.length; i++) { removed.push(p.store_sites[i][1]); p.code[p.store_sites[i][0]] = word_clear(p.code[p.store_sites[i][0]], STORE_FLAG); }
    removed.sort(function (a, b) { return a - b; });
    function shift(off) { var n = 0, k; for (k = 0; k < removed.length && removed[k] < off; k++) { n++; } return off - n; }
    for (i = 0; i < this.abs_words.length; i++) { p.code[this.abs_words[i]] = shift(p.code[this.abs_words[i]]); }
    for (i = 0; i < this.abs_words.length; i++) { self.abs_words[i] = shift(this.abs_words[i]); }
    for (i = removed.length - 1; i >= 0; i--) { p.code.splice(removed[i], 1); }
    p.store_sites = [];
    return der_off >= 0 ? shift(der_off) : der_off;
  };
  Lowering.prototype.stat_mode_applies = function () {
    var terms = this.terms, plates = [], t, i, points = 0;
    if (!this.stat_lowering) { return false; }
    for (t = 0; t < terms.length; t++) { if (terms[t].hasOwnProperty("plate_kind")) { plates.push(terms[t]); } }
    if (!plates.length || !this.param_ranges.length) { return false; }
    for (i = 0; i < this.param_ranges.length; i++) { if (this.param_ranges[i][2] === "binary") { return false; } }
    for (i = 0; i < plates.length; i++) { if (plates[i].plate_kind !== PLATE_NORM_IID || set_size(plates[i].mean_deps) !== 1) { return false; } points += plates[i].n; }
    for (t = 0; t < terms.length; t++) { if (terms[t].kind !== "value") { return false; } }
    if (terms.length + plates.length > MAX_IMMEDIATE) { return false; }
    return points >= MIN_STAT_POINTS;
  };
  Lowering.prototype.emit_stat_programs = function () {
    var p = this.prog, terms = this.terms, n_sum = terms.length, stats = [], t, c, tr, st, run_start, run_len, moved, k, frag, total = 0;
    for (t = 0; t < terms.length; t++) { if (terms[t].stat) { stats.push(terms[t].stat); } }
    for (k = 0; k < stats.length; k++) { stats[k].slot = n_sum + stats[k].k; p.code[stats[k].slot_word] = stats[k].slot; total += stats[k].n; }
    p.n_sum_terms = n_sum; p.n_terms = n_sum + stats.length;
    for (c = 0; c < this.n_comp; c++) {
      p.comp_prog.push(p.code.length);
      p.touch_off.push(p.touch_terms.length);
      run_start = null; run_len = 0;
      for (t = 0; t < terms.length; t++) {
        tr = terms[t];
        if (!tr.deps.hasOwnProperty(c)) {
          if (run_len && run_start + run_len === t) { run_len++; }
          else { if (run_len) { p.emit("ACC_RANGE", run_start, [run_len]); } run_start = t; run_len = 1; }
          continue;
        }
        if (run_len) { p.emit("ACC_RANGE", run_start, [run_len]); }
        run_start = null; run_len = 0;
        st = tr.stat;
        if (!st) { frag = p.code.slice(tr.start, tr.end); for (k = 0; k < frag.length; k++) { p.code.push(frag[k]); } }
        else {
          moved = st.mean_deps.hasOwnProperty(c);
          p.emit(moved ? "CAND" : "CACHED", st.slot);
          frag = p.code.slice(st.split, tr.end);
          for (k = 0; k < frag.length; k++) { p.code.push(frag[k]); }
          if (moved) { p.touch_terms.push(st.slot); }
        }
        p.touch_terms.push(t);
      }
      if (run_len) { p.emit("ACC_RANGE", run_start, [run_len]); }
      p.emit("END");
    }
    p.touch_off.push(p.touch_terms.length);
    p.stat_prog = p.code.length;
    for (t = 0; t < terms.length; t++) {
      st = terms[t].stat;
      if (st) {
        frag = p.code.slice(terms[t].start, st.split);
        frag[st.ss_word - terms[t].start] = word_or(frag[st.ss_word - terms[t].start], ACC_FLAG);
        for (k = 0; k < frag.length; k++) { p.code.push(frag[k]); }
      }
    }
    p.emit("END");
    p.summary.push("pre-evaluated statistics: " + stats.length + " plate(s), " + total + " points, one data pass per sweep");
  };
  Lowering.prototype.lower = function (result, derived_names, derived) {
    var p = this.prog, mark, off;
    this.record_terms = true;
    mark = [p.code.length, p.plates.length, p.summary.length, p.store_sites.length, this.abs_words.length];
    off = this.add_logpost(result, derived_names, derived);
    if (!this.faithful && this.stat_mode_applies()) {
      p.code.length = mark[0]; p.plates.length = mark[1]; p.summary.length = mark[2]; p.store_sites.length = mark[3]; this.abs_words.length = mark[4];
      this.terms = [];
      this.stat_mode = true;
      off = this.add_logpost(result, derived_names, derived);
      this.stat_mode = false;
      this.record_terms = false;
      this.emit_stat_programs();
      p.logpost_prog = off[0]; p.derived_prog = off[1];
      return this.finish();
    }
    this.record_terms = false;
    if (this.cache_worthwhile()) { this.emit_component_programs(); } else { off[1] = this.strip_stores(off[1]); }
    p.logpost_prog = off[0]; p.derived_prog = off[1];
    return this.finish();
  };

  // ------------------------------------------------------------------------------------------------ tracing
  function global_lookup(name) {
    try { return (new Function("return typeof " + name + " !== 'undefined' ? " + name + " : undefined;"))(); } catch (e) { return undefined; }
  }
  // instantiate the rewritten closure; free identifiers: ld / Math -> the recording versions, functions -> rewritten too
  function instantiate(fn, R, scope, depth) {
    var rw = rewriter.rewrite(fn.toString()), names = ["__r"], values = [R], i, name, v, k;
    for (i = 0; i < rw.free.length; i++) {
      name = rw.free[i];
      if (name === "ld") { v = R.ld; }
      else if (name === "Math") { v = R.Math; }
      else {
        v = (scope && scope.hasOwnProperty(name)) ? scope[name] : global_lookup(name);
        if (v === undefined) { throw "log_post refers to `" + name + "`, which is not visible to the sampler: pass it in options.scope"; }
        if (typeof v === "function" && depth < 8 && v.toString().indexOf("[native code]") < 0) { v = instantiate(v, R, scope, depth + 1); }
        else if (v && typeof v === "object" && typeof v.norm === "function" && typeof v.lgamma === "function") { v = R.ld; }      // the ld module under another name
      }
      names.push(name); values.push(v);
    }
    var decl = "";
    for (k in rw.implicit) { if (rw.implicit.hasOwnProperty(k)) { decl += "var " + k + ";\n"; } }          // sloppy-mode implicit globals stay local to the recording
    names.push(decl + "return " + rw.source + ";");
    return Function.apply(null, names).apply(null, values);
  }
  function run_closure(tr, fn, names, params, offsets, wrapped, base_state) {
    var state = tr.make_state(names, params, offsets), result, derived = {}, dnames = [], i, k, v;
    state.base = base_state || null;
    result = fn(state, wrapped);
    if (result === undefined) { throw "log_post returned undefined"; }
    result = lift(result);
    for (i = 0; i < state.keys.length; i++) {
      k = state.keys[i];
      if (!params.hasOwnProperty(k)) {
        v = state.vals[k];
        if (!is_sym(v) && !is_num(v)) { throw "derived quantity " + k + " must be a number"; }
        dnames.push(k); derived[k] = v;
      }
    }
    return [result, dnames, derived];
  }
  // -> program (the fields of amwg_model, see mcmc.js / amwg_napi.cc) ; options: {faithful, scope, stat_lowering}
  function trace(log_post, names, params, offsets, n_comp, data, options) {
    options = options || {};
    var tr = new Tracer(), wrapped = tr.wrap_data(data), ranges = [], i, p, n, R = make_runtime(), fn, low, out, comps = [], c, v, k, prog, off;
    for (i = 0; i < names.length; i++) {
      p = params[names[i]]; n = 1;
      for (k = 0; k < p.dim.length; k++) { n *= p.dim[k]; }
      ranges.push([offsets[names[i]], n, p.type]);
      if (p.type === "binary") { for (c = 0; c < n; c++) { comps.push(offsets[names[i]] + c); } }
    }
    fn = instantiate(log_post, R, options.scope || null, 0);
    low = new Lowering(tr, n_comp, options.faithful, ranges);
    if (options.stat_lowering === false) { low.stat_lowering = false; }
    try {
      out = run_closure(tr, fn, names, params, offsets, wrapped, options.base_state);
    } catch (exc) {
      if (!(exc instanceof NeedsConcrete)) { throw exc; }
      if (!comps.length || comps.length > MAX_VARIANT_COMPS) { throw exc.message; }
      prog = low.prog;
      prog.variant_comps = comps;
      for (v = 0; v < Math.pow(2, comps.length); v++) {
        tr.concrete = {};
        for (k = 0; k < comps.length; k++) { tr.concrete[comps[k]] = Math.floor(v / Math.pow(2, k)) % 2; }
        try { out = run_closure(tr, fn, names, params, offsets, wrapped, options.base_state); }
        catch (exc2) { if (exc2 instanceof NeedsConcrete) { throw exc2.message; } throw exc2; }
        off = low.add_logpost(out[0], out[1], out[2]);
        prog.variant_logpost.push(off[0]);
        prog.variant_derived.push(off[1]);
      }
      tr.concrete = {};
      prog.logpost_prog = prog.variant_logpost[0]; prog.derived_prog = prog.variant_derived[0];
      low.finish();
      return prog;
    }
    prog = low.lower(out[0], out[1], out[2]);
    prog.derived_names = out[1];
    return prog;
  }

  return {trace: trace, where: where, Sym: Sym, OP: OP, NeedsConcrete: NeedsConcrete};
}));
