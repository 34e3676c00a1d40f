// sde_dsl.cpp — the `@sde_model` front end, re-implemented in C++ so that the host language only has to
// hand over the macro body as text.  It computes what src/models/codegen.jl:87-262 computes:
//
//   equations  -> SDE lines (lhs a symbol), correlation lines (lhs `dW1*dW2`), mark lines (`~`)   codegen.jl:95-97
//   state vars from `d<name>` on the left, Wiener symbols `d[wW]...` and jump symbols `dN...` on the right,
//   in order of appearance                                                                         codegen.jl:99-103
//   drift_i     = factor of dt   (substitute dt=1, other differentials 0, simplify)               codegen.jl:106,178-183
//   diffusion   = [factor of dW_j] * chol(correlation matrix), symbolically                        codegen.jl:105,111-112,213-262
//   parameters  = symbols of drift ∪ diffusion minus state vars and t, in order of appearance     codegen.jl:116-118
//   kind        = JumpSDE | StateIndependentDiffusion | AbstractSDE                                codegen.jl:120-126
//
// and then emits a CUDA device functor (struct GenModel) with the interface of sde_models_builtin.cuh.
// The simplifier follows the rule set of Calculus.jl's `simplify` (Calculus 0.5.0, pinned in Manifest.toml:24-28):
// drop +0 / *1, x*0 -> 0, x-x -> 0, x/1 -> x, 0/x -> 0, x^1 -> x, x^0 -> 1, fold numeric arguments (sum / product
// moved to the front), iterate to a fixed point.  Operand order of what survives follows the source text, so the
// generated arithmetic has the rounding sequence of the Julia methods the macro would have generated.
#include "sde_dsl.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>

namespace {

struct Expr;
typedef std::shared_ptr<const Expr> E;
struct Expr {
  enum Kind { NUM, SYM, CALL } kind;
  double num = 0.0;
  bool is_int = false;
  std::string name;  // SYM: identifier; CALL: operator or function name
  std::vector<E> args;
};

E mk_num(double v, bool is_int) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::NUM;
  e->num = v;
  e->is_int = is_int;
  return e;
}
E mk_sym(const std::string &s) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::SYM;
  e->name = s;
  return e;
}
E mk_call(const std::string &f, std::vector<E> a) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::CALL;
  e->name = f;
  e->args = std::move(a);
  return e;
}
bool is_num(const E &e) { return e->kind == Expr::NUM; }
bool is_val(const E &e, double v) { return e->kind == Expr::NUM && e->num == v; }

bool equal(const E &a, const E &b) {
  if (a->kind != b->kind) return false;
  if (a->kind == Expr::NUM) return a->num == b->num;
  if (a->kind == Expr::SYM) return a->name == b->name;
  if (a->name != b->name || a->args.size() != b->args.size()) return false;
  for (size_t i = 0; i < a->args.size(); i++)
    if (!equal(a->args[i], b->args[i])) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// tokenizer / parser (Julia-like infix: n-ary + and *, left-assoc - and /, right-assoc ^, unary -,
// calls, numeric-literal juxtaposition such as 2θ)
// ------------------------------------------------------------------------------------------------
struct Tok {
  enum T { NUM, ID, OP, END } t;
  std::string s;
  double v = 0;
  bool is_int = false;
  bool glued = false;  // no whitespace between the previous token and this one
};

bool id_start(unsigned char c) { return isalpha(c) || c == '_' || c >= 0x80; }
bool id_char(unsigned char c) { return id_start(c) || isdigit(c); }

std::vector<Tok> tokenize(const std::string &s) {
  std::vector<Tok> out;
  size_t i = 0;
  bool glued = false;
  while (i < s.size()) {
    unsigned char c = s[i];
    if (c == ' ' || c == '\t' || c == '\r') { i++; glued = false; continue; }
    Tok t;
    t.glued = glued;
    if (isdigit(c) || (c == '.' && i + 1 < s.size() && isdigit((unsigned char)s[i + 1]))) {
      size_t j = i;
      bool isint = true;
      while (j < s.size() && isdigit((unsigned char)s[j])) j++;
      if (j < s.size() && s[j] == '.' && !(j + 1 < s.size() && (s[j + 1] == '*' || s[j + 1] == '/' || s[j + 1] == '^'))) {
        isint = false;
        j++;
        while (j < s.size() && isdigit((unsigned char)s[j])) j++;
      }
      if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
        size_t k = j + 1;
        if (k < s.size() && (s[k] == '+' || s[k] == '-')) k++;
        if (k < s.size() && isdigit((unsigned char)s[k])) {
          isint = false;
          while (k < s.size() && isdigit((unsigned char)s[k])) k++;
          j = k;
        }
      }
      t.t = Tok::NUM;
      t.s = s.substr(i, j - i);
      t.v = strtod(t.s.c_str(), nullptr);
      t.is_int = isint;
      i = j;
    } else if (id_start(c)) {
      size_t j = i;
      while (j < s.size() && id_char((unsigned char)s[j])) j++;
      t.t = Tok::ID;
      t.s = s.substr(i, j - i);
      i = j;
    } else {
      t.t = Tok::OP;
      t.s = std::string(1, (char)c);
      i++;
    }
    out.push_back(t);
    glued = true;
  }
  Tok e;
  e.t = Tok::END;
  out.push_back(e);
  return out;
}

struct Parser {
  std::vector<Tok> toks;
  size_t p = 0;
  explicit Parser(const std::string &s) : toks(tokenize(s)) {}
  const Tok &peek() const { return toks[p]; }
  bool is_op(const char *o) const { return toks[p].t == Tok::OP && toks[p].s == o; }
  void expect(const char *o) {
    if (!is_op(o)) throw std::runtime_error(std::string("expected '") + o + "' near '" + toks[p].s + "'");
    p++;
  }
  E parse_expr() { return parse_sum(); }
  // a + b - c + d parses in Julia as ((a + b) - c) + d, with runs of + collected into one n-ary call
  E parse_sum() {
    E lhs = parse_prod();
    while (is_op("+") || is_op("-")) {
      if (is_op("+")) {
        std::vector<E> a{lhs};
        while (is_op("+")) { p++; a.push_back(parse_prod()); }
        lhs = mk_call("+", a);
      } else {
        p++;
        E r = parse_prod();
        lhs = mk_call("-", {lhs, r});
      }
    }
    return lhs;
  }
  E parse_prod() {
    E lhs = parse_unary();
    while (is_op("*") || is_op("/")) {
      if (is_op("*")) {
        std::vector<E> a{lhs};
        while (is_op("*")) { p++; a.push_back(parse_unary()); }
        lhs = mk_call("*", a);
      } else {
        p++;
        E r = parse_unary();
        lhs = mk_call("/", {lhs, r});
      }
    }
    return lhs;
  }
  E parse_unary() {
    if (is_op("-")) {
      p++;
      E a = parse_unary();
      if (is_num(a)) return mk_num(-a->num, a->is_int);
      return mk_call("-", {a});
    }
    if (is_op("+")) { p++; return parse_unary(); }
    return parse_pow();
  }
  E parse_pow() {
    E base = parse_atom();
    if (is_op("^")) {
      p++;
      E ex = parse_unary();  // right associative, binds tighter than unary minus on the left
      return mk_call("^", {base, ex});
    }
    return base;
  }
  E parse_atom() {
    const Tok t = peek();
    if (t.t == Tok::NUM) {
      p++;
      E n = mk_num(t.v, t.is_int);
      // numeric-literal juxtaposition: 2θ, 2(x+y)
      if ((peek().t == Tok::ID || is_op("(")) && peek().glued) {
        E r = parse_pow();
        return mk_call("*", {n, r});
      }
      return n;
    }
    if (t.t == Tok::ID) {
      p++;
      if (is_op("(") && peek().glued) {
        p++;
        std::vector<E> a;
        if (!is_op(")")) {
          a.push_back(parse_expr());
          while (is_op(",")) { p++; a.push_back(parse_expr()); }
        }
        expect(")");
        return mk_call(t.s, a);
      }
      return mk_sym(t.s);
    }
    if (is_op("(")) {
      p++;
      E e = parse_expr();
      expect(")");
      return e;
    }
    throw std::runtime_error("unexpected token '" + t.s + "'");
  }
};

E parse_string(const std::string &s) {
  Parser ps(s);
  E e = ps.parse_expr();
  if (ps.peek().t != Tok::END) throw std::runtime_error("trailing input near '" + ps.peek().s + "' in '" + s + "'");
  return e;
}

// ------------------------------------------------------------------------------------------------
// printing (Julia-like, re-parseable)
// ------------------------------------------------------------------------------------------------
bool is_operator(const std::string &n) { return n == "+" || n == "-" || n == "*" || n == "/" || n == "^"; }
int prec_of(const E &e) {
  if (e->kind != Expr::CALL || !is_operator(e->name)) return 10;
  if (e->name == "+" || (e->name == "-" && e->args.size() == 2)) return 1;
  if (e->name == "*" || e->name == "/") return 2;
  if (e->name == "-") return 3;  // unary
  return 4;                      // ^
}
std::string num_str(double v, bool is_int) {
  char b[64];
  if (is_int && fabs(v) < 1e15) snprintf(b, sizeof b, "%lld", (long long)v);
  else {
    snprintf(b, sizeof b, "%.17g", v);
    if (!strpbrk(b, ".en")) strcat(b, ".0");
  }
  return b;
}
std::string to_str(const E &e) {
  if (e->kind == Expr::NUM) return num_str(e->num, e->is_int);
  if (e->kind == Expr::SYM) return e->name;
  if (!is_operator(e->name)) {
    std::string s = e->name + "(";
    for (size_t i = 0; i < e->args.size(); i++) s += (i ? ", " : "") + to_str(e->args[i]);
    return s + ")";
  }
  const int pr = prec_of(e);
  auto wrap = [&](const E &a, bool strict) {
    std::string s = to_str(a);
    const bool neg = is_num(a) && a->num < 0;
    if (neg || prec_of(a) < pr || (strict && prec_of(a) == pr)) return "(" + s + ")";
    return s;
  };
  if (e->name == "-" && e->args.size() == 1) return "-" + wrap(e->args[0], true);
  std::string s = wrap(e->args[0], e->name == "^");
  for (size_t i = 1; i < e->args.size(); i++) s += " " + e->name + " " + wrap(e->args[i], true);
  return s;
}

// ------------------------------------------------------------------------------------------------
// symbol collection in order of appearance (codegen.jl:203-211: `symbols` skips the callee of calls)
// ------------------------------------------------------------------------------------------------
void symbols(const E &e, std::vector<std::string> &out) {
  if (e->kind == Expr::SYM) {
    if (std::find(out.begin(), out.end(), e->name) == out.end()) out.push_back(e->name);
  } else if (e->kind == Expr::CALL) {
    for (auto &a : e->args) symbols(a, out);
  }
}

E substitute(const E &e, const std::map<std::string, E> &m) {
  if (e->kind == Expr::SYM) {
    auto it = m.find(e->name);
    return it == m.end() ? e : it->second;
  }
  if (e->kind == Expr::CALL) {
    std::vector<E> a;
    for (auto &x : e->args) a.push_back(substitute(x, m));
    return mk_call(e->name, a);
  }
  return e;
}

// ------------------------------------------------------------------------------------------------
// numeric evaluation of functions (constant folding and the host evaluator used by the tests)
// ------------------------------------------------------------------------------------------------
bool eval_fn(const std::string &f, const std::vector<double> &a, double *r) {
  const size_t n = a.size();
  if (f == "+") { double s = n ? a[0] : 0; for (size_t i = 1; i < n; i++) s = s + a[i]; *r = s; return true; }
  if (f == "*") { double s = n ? a[0] : 1; for (size_t i = 1; i < n; i++) s = s * a[i]; *r = s; return true; }
  if (f == "-") { if (n == 1) { *r = -a[0]; return true; } if (n == 2) { *r = a[0] - a[1]; return true; } return false; }
  if (f == "/" && n == 2) { *r = a[0] / a[1]; return true; }
  if (f == "^" && n == 2) {
    if (a[1] == 2) *r = a[0] * a[0];
    else if (a[1] == 3) *r = a[0] * a[0] * a[0];
    else *r = pow(a[0], a[1]);
    return true;
  }
  if (n == 1) {
    const double x = a[0];
    if (f == "sqrt") { *r = sqrt(x); return true; }
    if (f == "exp") { *r = exp(x); return true; }
    if (f == "log") { *r = log(x); return true; }
    if (f == "sin") { *r = sin(x); return true; }
    if (f == "cos") { *r = cos(x); return true; }
    if (f == "tan") { *r = tan(x); return true; }
    if (f == "tanh") { *r = tanh(x); return true; }
    if (f == "sinh") { *r = sinh(x); return true; }
    if (f == "cosh") { *r = cosh(x); return true; }
    if (f == "atan") { *r = atan(x); return true; }
    if (f == "abs") { *r = fabs(x); return true; }
    if (f == "log1p") { *r = log1p(x); return true; }
    if (f == "expm1") { *r = expm1(x); return true; }
    if (f == "sign") { *r = (x > 0) - (x < 0); return true; }
    if (f == "inv") { *r = 1.0 / x; return true; }
  }
  if (n == 2) {
    if (f == "max") { *r = fmax(a[0], a[1]); return true; }
    if (f == "min") { *r = fmin(a[0], a[1]); return true; }
    if (f == "atan") { *r = atan2(a[0], a[1]); return true; }
  }
  return false;
}

E fold(const E &e) {  // all arguments numeric -> evaluate (Calculus.simplify's `eval` branch)
  std::vector<double> a;
  bool all_int = true;
  for (auto &x : e->args) { a.push_back(x->num); all_int = all_int && x->is_int; }
  double r;
  if (!eval_fn(e->name, a, &r)) return e;
  const bool int_op = e->name == "+" || e->name == "*" || e->name == "-" || (e->name == "^" && a.size() == 2 && a[1] >= 0);
  return mk_num(r, all_int && int_op);
}

E simplify(const E &e);

E simplify_once(const E &e) {
  if (e->kind != Expr::CALL) return e;
  bool all_numeric = !e->args.empty();
  for (auto &a : e->args) all_numeric = all_numeric && is_num(a);
  if (all_numeric) {
    E f = fold(e);
    if (f != e) return f;
  }
  const std::string &op = e->name;
  std::vector<E> args;
  if (op == "+") {
    for (auto &a : e->args) if (!is_val(a, 0)) args.push_back(simplify(a));
    if (args.empty()) return mk_num(0, true);
    if (args.size() == 1) return args[0];
    double sum = 0;
    bool sint = true, any = false;
    std::vector<E> sym;
    for (auto &a : args) {
      if (is_num(a)) { sum += a->num; sint = sint && a->is_int; any = true; }
      else sym.push_back(a);
    }
    std::vector<E> out;
    if (any && sum != 0) out.push_back(mk_num(sum, sint));
    for (auto &s : sym) out.push_back(s);
    return mk_call("+", out);
  }
  if (op == "-") {
    for (auto &a : e->args) args.push_back(simplify(a));
    if (args.size() == 2 && equal(args[0], args[1])) return mk_num(0, true);
    if (args.size() == 2 && is_val(args[1], 0)) return args[0];
    if (args.size() == 2 && is_val(args[0], 0)) return mk_call("-", {args[1]});
    return mk_call("-", args);
  }
  if (op == "*") {
    for (auto &a : e->args) if (!is_val(a, 1)) args.push_back(simplify(a));
    if (args.empty()) return mk_num(1, true);
    if (args.size() == 1) return args[0];
    for (auto &a : args) if (is_val(a, 0)) return mk_num(0, true);
    double prod = 1;
    bool pint = true, any = false;
    std::vector<E> sym;
    for (auto &a : args) {
      if (is_num(a)) { prod *= a->num; pint = pint && a->is_int; any = true; }
      else sym.push_back(a);
    }
    std::vector<E> out;
    if (any && prod != 1) out.push_back(mk_num(prod, pint));
    for (auto &s : sym) out.push_back(s);
    return mk_call("*", out);
  }
  for (auto &a : e->args) args.push_back(simplify(a));
  if (op == "/" && args.size() == 2) {
    if (is_val(args[1], 1)) return args[0];
    if (is_val(args[0], 0)) return mk_num(0, true);
  }
  if (op == "^" && args.size() == 2) {
    if (is_val(args[1], 0)) return mk_num(1, true);
    if (is_val(args[1], 1)) return args[0];
  }
  return mk_call(op, args);
}

E simplify(const E &e) {
  E cur = e;
  for (int it = 0; it < 64; it++) {
    E nx = simplify_once(cur);
    if (equal(nx, cur)) return nx;
    cur = nx;
  }
  return cur;
}

// factor_extract(ex, one_sym, zero_syms)   (codegen.jl:178-183)
E factor_extract(const E &ex, const std::string &one_sym, const std::vector<std::string> &zero_syms) {
  std::map<std::string, E> m;
  for (auto &z : zero_syms) m[z] = mk_num(0, true);
  m[one_sym] = mk_num(1, true);
  return simplify(substitute(ex, m));
}

// ------------------------------------------------------------------------------------------------
// symbolic differentiation (the reference differentiates `diffusion` with ForwardDiff, milstein.jl:1-2;
// exact differentiation rules give the same derivative up to rounding)
// ------------------------------------------------------------------------------------------------
E diff(const E &e, const std::string &v) {
  if (e->kind == Expr::NUM) return mk_num(0, true);
  if (e->kind == Expr::SYM) return mk_num(e->name == v ? 1 : 0, true);
  const std::string &f = e->name;
  const auto &a = e->args;
  auto d = [&](size_t i) { return diff(a[i], v); };
  if (f == "+") { std::vector<E> t; for (size_t i = 0; i < a.size(); i++) t.push_back(d(i)); return mk_call("+", t); }
  if (f == "-") { std::vector<E> t; for (size_t i = 0; i < a.size(); i++) t.push_back(d(i)); return mk_call("-", t); }
  if (f == "*") {
    std::vector<E> terms;
    for (size_t i = 0; i < a.size(); i++) {
      std::vector<E> fac;
      for (size_t j = 0; j < a.size(); j++) fac.push_back(j == i ? d(i) : a[j]);
      terms.push_back(mk_call("*", fac));
    }
    return mk_call("+", terms);
  }
  if (f == "/" && a.size() == 2)
    return mk_call("/", {mk_call("-", {mk_call("*", {d(0), a[1]}), mk_call("*", {a[0], d(1)})}), mk_call("*", {a[1], a[1]})});
  if (f == "^" && a.size() == 2) {
    if (is_num(a[1]))
      return mk_call("*", {mk_num(a[1]->num, a[1]->is_int), mk_call("^", {a[0], mk_num(a[1]->num - 1, a[1]->is_int)}), d(0)});
    return mk_call("*", {e, mk_call("+", {mk_call("*", {d(1), mk_call("log", {a[0]})}),
                                         mk_call("/", {mk_call("*", {a[1], d(0)}), a[0]})})});
  }
  if (a.size() == 1) {
    const E u = a[0], du = d(0);
    if (f == "sqrt") return mk_call("/", {du, mk_call("*", {mk_num(2, true), e})});
    if (f == "exp") return mk_call("*", {e, du});
    if (f == "log") return mk_call("/", {du, u});
    if (f == "sin") return mk_call("*", {mk_call("cos", {u}), du});
    if (f == "cos") return mk_call("*", {mk_call("-", {mk_call("sin", {u})}), du});
    if (f == "tan") return mk_call("*", {mk_call("+", {mk_num(1, true), mk_call("^", {e, mk_num(2, true)})}), du});
    if (f == "tanh") return mk_call("*", {mk_call("-", {mk_num(1, true), mk_call("^", {e, mk_num(2, true)})}), du});
    if (f == "sinh") return mk_call("*", {mk_call("cosh", {u}), du});
    if (f == "cosh") return mk_call("*", {mk_call("sinh", {u}), du});
    if (f == "atan") return mk_call("/", {du, mk_call("+", {mk_num(1, true), mk_call("^", {u, mk_num(2, true)})})});
    if (f == "abs") return mk_call("*", {mk_call("sign", {u}), du});
    if (f == "log1p") return mk_call("/", {du, mk_call("+", {mk_num(1, true), u})});
    if (f == "expm1") return mk_call("*", {mk_call("exp", {u}), du});
    if (f == "inv") return mk_call("/", {mk_call("-", {du}), mk_call("*", {u, u})});
  }
  throw std::runtime_error("cannot differentiate '" + f + "' (needed for the Milstein scheme)");
}

// ------------------------------------------------------------------------------------------------
// symbolic Cholesky and matrix product (codegen.jl:213-237) — same recurrences, same simplifier
// ------------------------------------------------------------------------------------------------
typedef std::vector<std::vector<E>> Mat;

Mat symbolic_chol(const Mat &A) {
  const size_t n = A.size();
  Mat B(n, std::vector<E>(n, mk_num(0, true)));
  for (size_t k = 0; k < n; k++) {
    std::vector<E> sq{mk_num(0, true)};
    for (size_t j = 0; j < k; j++) sq.push_back(mk_call("^", {B[k][j], mk_num(2, true)}));
    E tmp = simplify(mk_call("-", {A[k][k], mk_call("+", sq)}));
    B[k][k] = is_val(tmp, 1) ? mk_num(1, true) : simplify(mk_call("sqrt", {tmp}));
    for (size_t i = 0; i < n; i++) {
      if (i == k) continue;
      std::vector<E> s{mk_num(0, true)};
      for (size_t j = 0; j < k; j++) s.push_back(mk_call("*", {B[i][j], B[k][j]}));
      B[i][k] = simplify(mk_call("/", {mk_call("-", {A[i][k], mk_call("+", s)}), B[k][k]}));
    }
  }
  return B;
}

Mat symbolic_mul(const Mat &A, const Mat &B) {
  const size_t n = A.size(), m = B.empty() ? 0 : B[0].size(), kk = B.size();
  Mat C(n, std::vector<E>(m));
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < m; j++) {
      std::vector<E> t;
      for (size_t k = 0; k < kk; k++) t.push_back(mk_call("*", {A[i][k], B[k][j]}));
      C[i][j] = simplify(mk_call("+", t));
    }
  return C;
}

// ------------------------------------------------------------------------------------------------
// CUDA emission
// ------------------------------------------------------------------------------------------------
struct Emit {
  std::map<std::string, std::string> sym;  // symbol -> C expression
  std::string T = "T";
  std::string lit(double v, bool is_int) const {
    char b[64];
    if (is_int && fabs(v) < 1e15) snprintf(b, sizeof b, "%lld", (long long)v);
    else snprintf(b, sizeof b, "%.17g", v);
    std::string s = b;
    if (!is_int && !strpbrk(b, ".en")) s += ".0";
    if (!is_int && strpbrk(b, "n")) s = v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
    return "(" + T + ")" + (v < 0 ? "(" + s + ")" : s);
  }
  std::string ex(const E &e) const {
    if (e->kind == Expr::NUM) return lit(e->num, e->is_int);
    if (e->kind == Expr::SYM) {
      auto it = sym.find(e->name);
      if (it == sym.end()) throw std::runtime_error("unbound symbol '" + e->name + "'");
      return it->second;
    }
    const std::string &f = e->name;
    const auto &a = e->args;
    if (f == "+" || f == "*") {  // n-ary, left to right: ((a op b) op c)
      std::string s = ex(a[0]);
      for (size_t i = 1; i < a.size(); i++) s = "(" + s + " " + f + " " + ex(a[i]) + ")";
      return s;
    }
    if (f == "-") return a.size() == 1 ? "(-" + ex(a[0]) + ")" : "(" + ex(a[0]) + " - " + ex(a[1]) + ")";
    if (f == "/") return "(" + ex(a[0]) + " / " + ex(a[1]) + ")";
    if (f == "^") {
      if (is_num(a[1]) && a[1]->is_int) {  // Base.literal_pow: x^2 = x*x, x^3 = x*x*x, x^-1 = inv(x)
        const long long n = (long long)a[1]->num;
        const std::string b = ex(a[0]);
        if (n == 2) return "(" + b + " * " + b + ")";
        if (n == 3) return "((" + b + " * " + b + ") * " + b + ")";
        if (n == -1) return "(" + lit(1, true) + " / " + b + ")";
      }
      return "pow(" + ex(a[0]) + ", " + ex(a[1]) + ")";
    }
    static const std::map<std::string, std::string> fn = {
        {"sqrt", "SDE_SQRT"}, {"exp", "exp"}, {"log", "log"}, {"sin", "sin"}, {"cos", "cos"}, {"tan", "tan"},
        {"tanh", "tanh"}, {"sinh", "sinh"}, {"cosh", "cosh"}, {"atan", "atan"}, {"abs", "fabs"},
        {"log1p", "log1p"}, {"expm1", "expm1"}, {"max", "fmax"}, {"min", "fmin"}};
    if (f == "inv" && a.size() == 1) return "(" + lit(1, true) + " / " + ex(a[0]) + ")";
    if (f == "sign" && a.size() == 1) {
      const std::string b = ex(a[0]);
      return "(" + T + ")((" + b + " > " + lit(0, true) + ") - (" + b + " < " + lit(0, true) + "))";
    }
    if (f == "atan" && a.size() == 2) return "atan2(" + ex(a[0]) + ", " + ex(a[1]) + ")";
    auto it = fn.find(f);
    if (it == fn.end()) throw std::runtime_error("function '" + f + "' is not supported in device code");
    std::string s = it->second + "(";
    for (size_t i = 0; i < a.size(); i++) s += (i ? ", " : "") + ex(a[i]);
    return s + ")";
  }
};

bool depends_on(const E &e, const std::set<std::string> &names) {
  if (e->kind == Expr::SYM) return names.count(e->name) > 0;
  if (e->kind == Expr::CALL)
    for (auto &a : e->args)
      if (depends_on(a, names)) return true;
  return false;
}

// replace maximal parameter-only sub-expressions (no state variable, no t) by hoisted constants
E hoist(const E &e, const std::set<std::string> &dyn, std::vector<E> &hoisted) {
  if (e->kind != Expr::CALL) return e;
  if (!depends_on(e, dyn)) {
    for (size_t i = 0; i < hoisted.size(); i++)
      if (equal(hoisted[i], e)) return mk_sym("__h" + std::to_string(i));
    hoisted.push_back(e);
    return mk_sym("__h" + std::to_string(hoisted.size() - 1));
  }
  std::vector<E> a;
  for (auto &x : e->args) a.push_back(hoist(x, dyn, hoisted));
  return mk_call(e->name, a);
}

std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// sde_dsl_build
// ------------------------------------------------------------------------------------------------
bool sde_dsl_build(const std::string &name, const std::string &eqs, const std::vector<std::string> &param_names,
                   sde_dsl_model *out, std::string *err) {
  try {
    // split into equations (newline or ';'), drop comments and begin/end
    std::vector<std::string> lines;
    {
      std::string cur;
      for (char c : eqs + "\n") {
        if (c == '\n' || c == ';') {
          std::string l = cur;
          size_t h = l.find('#');
          if (h != std::string::npos) l = l.substr(0, h);
          l = trim(l);
          if (!l.empty() && l != "begin" && l != "end") lines.push_back(l);
          cur.clear();
        } else cur += c;
      }
    }
    struct Eq { E lhs, rhs; };
    std::vector<Eq> sde_eqs, corr_eqs;
    std::vector<std::string> marks;
    for (auto &l : lines) {
      size_t tpos = l.find('~');
      if (tpos != std::string::npos) {  // mark equation  ξ ~ Dist(...)   (codegen.jl:97)
        marks.push_back(trim(l.substr(0, tpos)));
        continue;
      }
      size_t epos = l.find('=');
      if (epos == std::string::npos) throw std::runtime_error("not an equation: '" + l + "'");
      Eq q{parse_string(l.substr(0, epos)), parse_string(l.substr(epos + 1))};
      if (q.lhs->kind == Expr::SYM) sde_eqs.push_back(q);
      else corr_eqs.push_back(q);
    }
    if (sde_eqs.empty()) throw std::runtime_error("no SDE equation `dX = ...` found");

    // model variables: (?<=^d).* on the lhs   (codegen.jl:99)
    std::vector<std::string> vars, lhs_syms;
    for (auto &q : sde_eqs) {
      const std::string &s = q.lhs->name;
      if (s.size() < 2 || s[0] != 'd') throw std::runtime_error("left-hand side '" + s + "' is not of the form d<variable>");
      lhs_syms.push_back(s);
      vars.push_back(s.substr(1));
    }
    // process variables ^d[wW].* and jump variables ^dN on the rhs, in order of appearance   (codegen.jl:101-102)
    std::vector<std::string> wien, jumps;
    for (auto &q : sde_eqs) {
      std::vector<std::string> ss;
      symbols(q.rhs, ss);
      for (auto &s : ss) {
        if (s.size() >= 2 && s[0] == 'd' && (s[1] == 'w' || s[1] == 'W')) {
          if (std::find(wien.begin(), wien.end(), s) == wien.end()) wien.push_back(s);
        } else if (s.size() >= 2 && s[0] == 'd' && s[1] == 'N') {
          if (std::find(jumps.begin(), jumps.end(), s) == jumps.end()) jumps.push_back(s);
        }
      }
    }
    std::vector<std::string> differentials{"dt"};
    for (auto &w : wien) differentials.push_back(w);
    for (auto &j : jumps) differentials.push_back(j);
    const int D = (int)sde_eqs.size(), M = (int)wien.size();

    // correlation matrix (codegen.jl:243-262) and its symbolic Cholesky factor
    Mat S(M, std::vector<E>(M));
    for (int i = 0; i < M; i++)
      for (int j = 0; j < M; j++) S[i][j] = mk_num(i == j ? 1 : 0, true);
    for (auto &q : corr_eqs) {
      if (q.lhs->kind != Expr::CALL || q.lhs->name != "*" || q.lhs->args.size() != 2 ||
          q.lhs->args[0]->kind != Expr::SYM || q.lhs->args[1]->kind != Expr::SYM)
        throw std::runtime_error("correlation equation must look like dW1*dW2 = ρ*dt, got '" + to_str(q.lhs) + "'");
      auto fi = std::find(wien.begin(), wien.end(), q.lhs->args[0]->name);
      auto fj = std::find(wien.begin(), wien.end(), q.lhs->args[1]->name);
      if (fi == wien.end() || fj == wien.end())
        throw std::runtime_error("correlation equation names an unknown Wiener process: '" + to_str(q.lhs) + "'");
      E rho = simplify(factor_extract(q.rhs, "dt", {}));
      S[fi - wien.begin()][fj - wien.begin()] = rho;
      S[fj - wien.begin()][fi - wien.begin()] = rho;
    }
    Mat chol = symbolic_chol(S);

    // drift and diffusion
    std::vector<E> drift;
    for (auto &q : sde_eqs) drift.push_back(factor_extract(q.rhs, "dt", differentials));
    Mat raw(D, std::vector<E>(M));
    for (int i = 0; i < D; i++)
      for (int j = 0; j < M; j++) raw[i][j] = factor_extract(sde_eqs[i].rhs, wien[j], differentials);
    Mat diffm = M > 0 ? symbolic_mul(raw, chol) : Mat(D, std::vector<E>());

    // parameters in order of appearance: drift vector, then the diffusion matrix column-major (SMatrix ctor order)
    std::vector<std::string> params = param_names;
    std::vector<std::string> diff_syms;
    for (int j = 0; j < M; j++)
      for (int i = 0; i < D; i++) symbols(diffm[i][j], diff_syms);
    if (params.empty()) {
      std::vector<std::string> all;
      for (auto &d : drift) symbols(d, all);
      for (auto &s : diff_syms) if (std::find(all.begin(), all.end(), s) == all.end()) all.push_back(s);
      for (auto &s : all) {
        const bool is_var = std::find(vars.begin(), vars.end(), s) != vars.end();
        const bool is_mark = std::find(marks.begin(), marks.end(), s) != marks.end();
        if (!is_var && s != "t" && !is_mark) params.push_back(s);
      }
    }
    int kind = 0;
    if (!jumps.empty()) kind = 2;
    else {
      bool state_dep = false;
      for (auto &s : diff_syms) state_dep = state_dep || std::find(vars.begin(), vars.end(), s) != vars.end();
      if (!state_dep) kind = 1;
    }

    out->D = D;
    out->M = M;
    out->kind = kind;
    out->params = params;
    out->vars = vars;
    out->wieners = wien;
    out->drift_str.clear();
    out->diffusion_str.clear();
    for (auto &d : drift) out->drift_str.push_back(to_str(d));
    if (M == 0) for (int i = 0; i < D; i++) out->diffusion_str.push_back("0");
    for (int i = 0; i < D; i++)
      for (int j = 0; j < M; j++) out->diffusion_str.push_back(to_str(diffm[i][j]));
    if (kind == 2) { out->cuda_source.clear(); return true; }
    if (D > 4 || M > 4 || params.size() > 16) { out->cuda_source.clear(); return true; }

    // every free symbol must be bound
    {
      std::vector<std::string> all;
      for (auto &d : drift) symbols(d, all);
      for (auto &s : diff_syms) all.push_back(s);
      for (auto &s : all) {
        const bool ok = std::find(vars.begin(), vars.end(), s) != vars.end() || s == "t" ||
                        std::find(params.begin(), params.end(), s) != params.end();
        if (!ok) throw std::runtime_error("symbol '" + s + "' is neither a state variable, t, nor a listed parameter");
      }
    }

    // ---- CUDA functor ----
    std::set<std::string> dyn(vars.begin(), vars.end());
    dyn.insert("t");
    std::vector<E> hoisted;
    std::vector<E> hdrift;
    Mat hdiff(D, std::vector<E>(M));
    for (auto &d : drift) hdrift.push_back(hoist(d, dyn, hoisted));
    for (int i = 0; i < D; i++)
      for (int j = 0; j < M; j++) hdiff[i][j] = hoist(diffm[i][j], dyn, hoisted);
    // Milstein (M == 1): (∂σ σ)_i = Σ_k ∂σ_i/∂x_k σ_k
    Mat jac;
    bool have_mil = (M == 1);
    if (have_mil) {
      try {
        jac.assign(D, std::vector<E>(D));
        for (int i = 0; i < D; i++)
          for (int k = 0; k < D; k++) jac[i][k] = hoist(simplify(diff(diffm[i][0], vars[k])), dyn, hoisted);
      } catch (const std::exception &e) {
        // no derivative (max, min, sign, two-argument functions ...): Milstein(Δt) must FAIL for this model, not run plain
        // Euler–Maruyama silently (the reference differentiates with ForwardDiff and would succeed); the caller records it
        have_mil = false;
        out->milstein_error = e.what();
      }
    }
    // Polar form of the step (M == 2): when every non-zero diffusion entry is a product that carries the same factor
    // sqrt(A) with A depending on the state (Heston-type models: sqrt(V) in every noise term), the FP64 kernels hand the
    // model a pair as (w, ca, sb) with dW = sqrt(w) (ca, sb), and the step takes ONE square root of A*w instead of the
    // radius' and the model's (sde_device.cuh: Noise<>, PolarOf<>; the hand-written Heston functor does the same).
    bool have_polar = false;
    E polar_arg, polar_raw;
    if (M == 2) {
      std::function<void(const E &, std::vector<E> &)> factors = [&](const E &e, std::vector<E> &f) {
        if (e->kind == Expr::CALL && e->name == "*") for (auto &a : e->args) factors(a, f);
        else f.push_back(e);
      };
      std::vector<std::vector<E>> fl;   // factor lists of the non-zero entries, row-major
      for (int i = 0; i < D; i++)
        for (int j = 0; j < M; j++)
          if (!is_val(diffm[i][j], 0)) { fl.emplace_back(); factors(diffm[i][j], fl.back()); }
      E cand;
      if (!fl.empty())
        for (auto &f : fl[0]) {
          if (!(f->kind == Expr::CALL && f->name == "sqrt" && f->args.size() == 1 && depends_on(f->args[0], dyn))) continue;
          bool all = true;
          for (auto &g : fl) {
            bool found = false;
            for (auto &h : g) found = found || equal(h, f);
            all = all && found;
          }
          if (all) { cand = f; break; }
        }
      if (cand) {
        have_polar = true;
        polar_raw = cand->args[0];
        polar_arg = hoist(cand->args[0], dyn, hoisted);
      }
    }
    out->polar = have_polar;
    // FAST-math forms of the Euler–Maruyama update (SDE_STRICT = 0 only; STRICT keeps the reference's literal operation
    // order).  Same algebra, fewer operations: (1) every term — drift_i*dt, sigma_ij*dw_j — is a product whose
    // parameter-only factors, Δt included, are folded into ONE constant computed in prepare(); (2) when every term of row i
    // carries the state x_i itself (geometric rows: r*S*dt + sigma*S*dW), x_i is factored out: x_i + x_i*A, one FMA.
    // The same for the polar step, whose noise of row i is G * (rest_i0*ca + rest_i1*sb).
    std::set<std::string> dyn2 = dyn;   // symbols that vary per step: states, t and the noise of the step
    for (const char *n : {"__dw0", "__dw1", "__dw2", "__dw3", "__ca", "__sb", "__G"}) dyn2.insert(n);
    auto flat = [&](const E &e) {
      std::vector<E> f;
      std::function<void(const E &)> go = [&](const E &x) {
        if (x->kind == Expr::CALL && x->name == "*") for (auto &a : x->args) go(a);
        else f.push_back(x);
      };
      go(e);
      return f;
    };
    auto product = [&](const std::vector<E> &f) { return f.empty() ? mk_num(1, true) : (f.size() == 1 ? f[0] : mk_call("*", f)); };
    // term -> c * (dynamic factors), the constant hoisted (dt is visible in prepare)
    auto fold_term = [&](const std::vector<E> &f) {
      std::vector<E> st, dy;
      for (auto &x : f) (depends_on(x, dyn2) ? dy : st).push_back(x);
      E c = simplify(product(st));
      if (c->kind == Expr::CALL || (c->kind == Expr::SYM && c->name == "dt")) c = hoist(c, dyn2, hoisted);
      if (dy.empty()) return c;
      if (is_val(c, 1)) return product(dy);
      std::vector<E> all{c};
      for (auto &x : dy) all.push_back(x);
      return mk_call("*", all);
    };
    // row i:  x_i + sum(terms)  ->  x_i * (c' + A) + (cB + B)
    //   terms whose only dynamic factor is affine in x_i with parameter-only other parts (a - x_i, x_i + a, ...: mean
    //   reversion) are expanded first; then c' = 1 + (sum of parameter-only coefficients of x_i), A = the coefficients of
    //   x_i that vary per step, cB = the sum of the parameter-only terms, B = the remaining terms.  Geometric rows come out
    //   as x*A + x, mean-reverting ones as x*c' + (cB + noise): one FMA each.
    //   The expansion of mean reversion is for Float64 only (`expand`): in FP32, 1 - κΔt rounded to 24 bits carries κΔt to
    //   only 2^-24 / (κΔt) relative and would shift the reversion level by that much at every step; FP32 keeps x + c (a - x).
    auto fast_row = [&](int i, const std::vector<std::vector<E>> &terms, bool expand) {
      const E xi = mk_sym(vars[i]);
      auto dynamic = [&](const E &e) { return depends_on(e, dyn2); };
      std::vector<std::vector<E>> ex;
      for (auto &f : terms) {
        int idx = -1, ndyn = 0;
        for (size_t k = 0; k < f.size(); k++)
          if (dynamic(f[k])) { ndyn++; idx = (int)k; }
        bool done = false;
        if (expand && ndyn == 1 && f[idx]->kind == Expr::CALL && (f[idx]->name == "+" || (f[idx]->name == "-" && f[idx]->args.size() == 2))) {
          const E &g = f[idx];
          std::vector<E> st;
          int sx = 0;
          bool ok = true;
          for (size_t q = 0; q < g->args.size(); q++) {
            const int sign = (g->name == "-" && q == 1) ? -1 : 1;
            if (equal(g->args[q], xi)) { ok = ok && sx == 0; sx = sign; }
            else if (!dynamic(g->args[q])) st.push_back(sign > 0 ? g->args[q] : mk_call("-", {g->args[q]}));
            else ok = false;
          }
          if (ok && sx != 0) {
            std::vector<E> others;
            for (size_t k = 0; k < f.size(); k++)
              if ((int)k != idx) others.push_back(f[k]);
            if (!st.empty()) {
              auto t = others;
              t.push_back(st.size() == 1 ? st[0] : mk_call("+", st));
              ex.push_back(t);
            }
            auto t2 = others;
            if (sx < 0) t2.push_back(mk_num(-1, true));
            t2.push_back(xi);
            ex.push_back(t2);
            done = true;
          }
        }
        if (!done) ex.push_back(f);
      }
      std::vector<E> Kp{mk_num(1, true)}, Ap, Bs, Bd;
      for (auto f : ex) {
        int kx = -1;
        for (size_t k = 0; k < f.size() && kx < 0; k++)
          if (equal(f[k], xi)) kx = (int)k;
        if (kx >= 0) f.erase(f.begin() + kx);
        bool dyn_rest = false;
        for (auto &x : f) dyn_rest = dyn_rest || dynamic(x);
        if (kx >= 0) (dyn_rest ? Ap : Kp).push_back(dyn_rest ? fold_term(f) : simplify(product(f)));
        else (dyn_rest ? Bd : Bs).push_back(dyn_rest ? fold_term(f) : simplify(product(f)));
      }
      auto constant = [&](const std::vector<E> &parts) {   // the sum of parameter-only parts as one prepared constant
        E c = simplify(parts.size() == 1 ? parts[0] : mk_call("+", parts));
        return (c->kind == Expr::CALL || (c->kind == Expr::SYM && c->name == "dt")) ? hoist(c, dyn2, hoisted) : c;
      };
      auto sum_of = [&](std::vector<E> v) { return v.size() == 1 ? v[0] : mk_call("+", v); };
      E core;
      if (Ap.empty()) core = Kp.size() == 1 ? xi : mk_call("*", {xi, constant(Kp)});   // x * (1 + K): mean reversion, growth
      else {   // x*(K + A) + x: one FMA after A, and the small increment is not rounded against 1
        std::vector<E> in;
        if (Kp.size() > 1) in.push_back(constant(std::vector<E>(Kp.begin() + 1, Kp.end())));
        for (auto &x : Ap) in.push_back(x);
        core = mk_call("+", {mk_call("*", {xi, sum_of(in)}), xi});
      }
      std::vector<E> rest;
      if (!Bs.empty()) rest.push_back(constant(Bs));
      for (auto &x : Bd) rest.push_back(x);
      if (rest.empty()) return core;
      return mk_call("+", {core, sum_of(rest)});
    };
    std::vector<E> fast_em(D), fast_polar(D), fast_em32(D), fast_polar32(D);
    for (int i = 0; i < D; i++) {
      std::vector<std::vector<E>> terms;
      if (!is_val(drift[i], 0)) { auto f = flat(drift[i]); f.push_back(mk_sym("dt")); terms.push_back(f); }
      for (int j = 0; j < M; j++)
        if (!is_val(diffm[i][j], 0)) { auto f = flat(diffm[i][j]); f.push_back(mk_sym("__dw" + std::to_string(j))); terms.push_back(f); }
      fast_em[i] = fast_row(i, terms, true);
      fast_em32[i] = fast_row(i, terms, false);
      if (have_polar) {
        // the noise of the row as ONE term G * N_i; x_i is common only if it is a factor of every rest_ij
        const E xi = mk_sym(vars[i]);
        std::vector<std::vector<E>> pt;
        if (!is_val(drift[i], 0)) { auto f = flat(drift[i]); f.push_back(mk_sym("dt")); pt.push_back(f); }
        std::vector<std::vector<E>> rest;
        for (int j = 0; j < M; j++)
          if (!is_val(diffm[i][j], 0)) {
            // un-hoisted rest: the entry's factors without one occurrence of the common root
            auto f = flat(diffm[i][j]);
            for (size_t k = 0; k < f.size(); k++)
              if (f[k]->kind == Expr::CALL && f[k]->name == "sqrt" && equal(f[k]->args[0], polar_raw)) { f.erase(f.begin() + k); break; }
            f.push_back(mk_sym(j == 0 ? "__ca" : "__sb"));
            rest.push_back(f);
          }
        if (!rest.empty()) {
          // x_i is taken out of the noise term only if it is a factor of every rest_ij (then G * N_i * x_i is an x_i-term)
          bool common = true;
          for (auto &f : rest) {
            bool has = false;
            for (auto &x : f) has = has || equal(x, xi);
            common = common && has;
          }
          std::vector<E> ns;
          for (auto f : rest) {
            if (common)
              for (size_t k = 0; k < f.size(); k++)
                if (equal(f[k], xi)) { f.erase(f.begin() + k); break; }
            ns.push_back(fold_term(f));
          }
          std::vector<E> nf{mk_sym("__G"), ns.size() == 1 ? ns[0] : mk_call("+", ns)};
          if (common) nf.push_back(xi);
          pt.push_back(nf);
        }
        fast_polar[i] = fast_row(i, pt, true);
        fast_polar32[i] = fast_row(i, pt, false);
      }
    }
    const int NP = (int)params.size();
    if (NP + (int)hoisted.size() > 32) throw std::runtime_error("too many hoisted constants");

    Emit em;
    for (int i = 0; i < NP; i++) em.sym[params[i]] = "c[" + std::to_string(i) + "]";
    for (size_t i = 0; i < hoisted.size(); i++) em.sym["__h" + std::to_string(i)] = "c[" + std::to_string(NP + (int)i) + "]";
    for (int i = 0; i < D; i++) em.sym[vars[i]] = "x" + std::to_string(i);
    em.sym["t"] = "t";

    std::ostringstream o;
    o << "// generated by sdeb200 from @sde_model " << name << "\n";
    for (auto &l : lines) o << "//   " << l << "\n";
    o << "namespace sdeb200 {\nstruct GenModel {\n";
    o << "  static constexpr int D = " << D << ", M = " << M << ", NP = " << NP << ";\n";
    o << "  template <typename T> SDE_HD static void prepare(const T *p, T dt, T *c) {\n    (void)dt; (void)p;\n";
    for (int i = 0; i < NP; i++) o << "    c[" << i << "] = p[" << i << "];  // " << params[i] << "\n";
    {
      Emit ep = em;  // hoisted expressions only see parameters and dt (and earlier hoisted constants do not nest)
      ep.sym["dt"] = "dt";
      for (size_t i = 0; i < hoisted.size(); i++)
        o << "    c[" << NP + (int)i << "] = " << ep.ex(hoisted[i]) << ";  // " << to_str(hoisted[i]) << "\n";
    }
    o << "  }\n";
    if (M == 1) {  // diffusion column at an arbitrary state: RungeKutta evaluates it at the predictor (runge_kutta.jl:8)
      Emit es = em;
      for (int i = 0; i < D; i++) es.sym[vars[i]] = "xx[" + std::to_string(i) + "]";
      o << "  template <typename T> SDE_HD static void sigma_col(const T *c, T t, const T *xx, T *s) {\n    (void)c; (void)t; (void)xx;\n";
      for (int i = 0; i < D; i++) o << "    s[" << i << "] = " << es.ex(hdiff[i][0]) << ";\n";
      o << "  }\n";
    }
    Emit ef = em;   // FAST forms: the step's noise symbols
    ef.sym["dt"] = "dt";
    for (int j = 0; j < 4; j++) ef.sym["__dw" + std::to_string(j)] = "dw[" + std::to_string(j) + "]";
    ef.sym["__ca"] = "ca";
    ef.sym["__sb"] = "sb";
    ef.sym["__G"] = "G";
    o << "  template <int SCHEME, typename T> SDE_HD static void step(const T *c, T t, T dt, const T *dw, T *x) {\n";
    o << "    (void)c; (void)t; (void)dw;\n";
    for (int i = 0; i < D; i++) o << "    const T x" << i << " = x[" << i << "];  // " << vars[i] << "\n";
    o << "#if !SDE_STRICT\n    if (SCHEME == 0) {  // FAST math: folded constants, common state factor (same algebra as below)\n";
    o << "      if (sizeof(T) == 8) {\n";
    for (int i = 0; i < D; i++) o << "        x[" << i << "] = " << ef.ex(fast_em[i]) << ";\n";
    o << "      } else {  // FP32: mean reversion stays x + c (a - x)\n";
    for (int i = 0; i < D; i++) o << "        x[" << i << "] = " << ef.ex(fast_em32[i]) << ";\n";
    o << "      }\n      return;\n    }\n#endif\n";
    for (int i = 0; i < D; i++) o << "    const T mu" << i << " = " << em.ex(hdrift[i]) << ";\n";
    for (int i = 0; i < D; i++)
      for (int j = 0; j < M; j++)
        if (!is_val(diffm[i][j], 0)) o << "    const T s" << i << "_" << j << " = " << em.ex(hdiff[i][j]) << ";\n";
    // x0 + μ*Δt + σ*Δw  ==  (x0 + μΔt) + (σ_i1 Δw_1 + σ_i2 Δw_2 + ...)   euler_maruyama.jl:4; structural zeros dropped
    for (int i = 0; i < D; i++) {
      std::string noise;
      for (int j = 0; j < M; j++) {
        if (is_val(diffm[i][j], 0)) continue;
        std::string term = "s" + std::to_string(i) + "_" + std::to_string(j) + " * dw[" + std::to_string(j) + "]";
        noise = noise.empty() ? term : "(" + noise + " + " + term + ")";
      }
      o << "    T y" << i << " = (x" << i << " + mu" << i << " * dt)";
      if (!noise.empty()) o << " + " << (noise[0] == '(' ? noise : "(" + noise + ")");
      o << ";\n";
    }
    if (have_mil) {
      // + ∂σ*σ*(Δw.^2 - Δt)/2   milstein.jl:8
      o << "    if (SCHEME == 1) {\n      const T cc = dw[0] * dw[0] - dt;\n";
      for (int i = 0; i < D; i++) {
        std::string js;
        for (int k = 0; k < D; k++) {
          if (is_val(jac[i][k], 0) && D > 1) continue;
          const std::string sk = is_val(diffm[k][0], 0) ? em.lit(0, true) : "s" + std::to_string(k) + "_0";
          std::string term = em.ex(jac[i][k]) + " * " + sk;
          js = js.empty() ? term : "(" + js + " + " + term + ")";
        }
        if (js.empty()) js = em.lit(0, true);
        o << "      y" << i << " = y" << i << " + ((" << js << ") * cc) / " << em.lit(2, true) << ";\n";
      }
      o << "    }\n";
    }
    if (M == 1) {
      // x0hat = x0 + μΔt + σ√Δt; σpred = diffusion(x0hat); + (σpred - σ)(Δw² - Δt)/(2√Δt)   runge_kutta.jl:7-9
      o << "    if (SCHEME == 2) {\n      const T sq = sqrt(dt);\n      T xh[" << D << "], sp[" << D << "];\n";
      for (int i = 0; i < D; i++) {
        const std::string si = is_val(diffm[i][0], 0) ? em.lit(0, true) : "s" + std::to_string(i) + "_0";
        o << "      xh[" << i << "] = (x" << i << " + mu" << i << " * dt) + " << si << " * sq;\n";
      }
      o << "      sigma_col(c, t, xh, sp);\n      const T cc = dw[0] * dw[0] - dt, den = " << em.lit(2, true) << " * sq;\n";
      for (int i = 0; i < D; i++) {
        const std::string si = is_val(diffm[i][0], 0) ? em.lit(0, true) : "s" + std::to_string(i) + "_0";
        o << "      y" << i << " = y" << i << " + ((sp[" << i << "] - " << si << ") * cc) / den;\n";
      }
      o << "    }\n";
    }
    for (int i = 0; i < D; i++) o << "    x[" << i << "] = y" << i << ";\n";
    o << "  }\n";
    if (have_polar) {
      o << "  static constexpr int POLAR = 1;\n";
      o << "  template <typename T> SDE_HD static void step_polar(const T *c, T t, T dt, T w, T ca, T sb, T *x) {\n";
      o << "    (void)c; (void)t; (void)dt;\n";
      for (int i = 0; i < D; i++) o << "    const T x" << i << " = x[" << i << "];  // " << vars[i] << "\n";
      o << "    const T G = SDE_SQRT(" << em.ex(polar_arg) << " * w);\n";
      o << "    if (sizeof(T) == 8) {\n";
      for (int i = 0; i < D; i++) o << "      x[" << i << "] = " << ef.ex(fast_polar[i]) << ";\n";
      o << "    } else {\n";
      for (int i = 0; i < D; i++) o << "      x[" << i << "] = " << ef.ex(fast_polar32[i]) << ";\n";
      o << "    }\n";
      o << "  }\n";
    }
    // drift / diffusion at a point, double precision, un-hoisted
    {
      Emit ec;
      ec.T = "double";
      for (int i = 0; i < NP; i++) ec.sym[params[i]] = "p[" + std::to_string(i) + "]";
      for (int i = 0; i < D; i++) ec.sym[vars[i]] = "x[" + std::to_string(i) + "]";
      ec.sym["t"] = "t";
      o << "  SDE_HD static void coefficients(const double *p, double t, const double *x, double *mu, double *sig) {\n";
      o << "    (void)p; (void)t; (void)x;\n";
      for (int i = 0; i < D; i++) o << "    mu[" << i << "] = " << ec.ex(drift[i]) << ";\n";
      for (int i = 0; i < D; i++)
        for (int j = 0; j < M; j++) o << "    sig[" << i << " * SDE_MAXM + " << j << "] = " << ec.ex(diffm[i][j]) << ";\n";
      if (M == 0) for (int i = 0; i < D; i++) o << "    sig[" << i << " * SDE_MAXM] = 0.0;\n";
      o << "  }\n";
    }
    o << "};\n}  // namespace sdeb200\n";
    out->cuda_source = o.str();
    return true;
  } catch (const std::exception &e) {
    *err = std::string("@sde_model ") + name + ": " + e.what();
    return false;
  }
}

bool sde_dsl_eval(const std::string &expr, const std::vector<std::string> &names, const std::vector<double> &values,
                  double *out, std::string *err) {
  try {
    E e = parse_string(expr);
    struct Ev {
      const std::vector<std::string> &n;
      const std::vector<double> &v;
      double go(const E &e) const {
        if (e->kind == Expr::NUM) return e->num;
        if (e->kind == Expr::SYM) {
          for (size_t i = 0; i < n.size(); i++) if (n[i] == e->name) return v[i];
          throw std::runtime_error("unbound symbol '" + e->name + "'");
        }
        std::vector<double> a;
        for (auto &x : e->args) a.push_back(go(x));
        double r;
        if (!eval_fn(e->name, a, &r)) throw std::runtime_error("cannot evaluate '" + e->name + "'");
        return r;
      }
    } ev{names, values};
    *out = ev.go(e);
    return true;
  } catch (const std::exception &e) {
    *err = e.what();
    return false;
  }
}

// C entry points for the CPU unit tests of the front end (no device involved)
extern "C" int sdeb200_debug_eval(const char *expr, const char *const *names, const double *values, int n, double *out) {
  std::vector<std::string> nn;
  std::vector<double> vv;
  for (int i = 0; i < n; i++) { nn.push_back(names[i]); vv.push_back(values[i]); }
  std::string err;
  return sde_dsl_eval(expr, nn, vv, out, &err) ? 0 : -2;
}
extern "C" int sdeb200_debug_simplify(const char *expr, char *buf, int cap) {
  try {
    std::string s = to_str(simplify(parse_string(expr)));
    snprintf(buf, cap, "%s", s.c_str());
    return 0;
  } catch (const std::exception &e) {
    snprintf(buf, cap, "%s", e.what());
    return -2;
  }
}
