// dtwa_expr.hpp -- host-side compiler from an observable expression string to the postfix bytecode
// the kernels interpret (dtwa_device.cuh VmOp).
//
// Replaces TWA.function_from_expression (src/TWA.jl:277-299), which turns a Symbol / Expr /
// SymEngine object into a Julia lambda through toString + Meta.parse + eval, and validates that the
// free symbols are a subset of the system's varnames (src/TWA.jl:291-293).  Grammar (Julia-like):
//   expr  := term (('+'|'-') term)*
//   term  := unary (('*'|'/') unary)*          also juxtaposition  2Q, 2(Q+1)
//   unary := ('-'|'+') unary | power
//   power := atom ('^' unary)?                 right associative, -Q^2 == -(Q^2); '**' == '^'
//   atom  := number | pi | name | func '(' expr (',' expr)* ')' | '(' expr ')'
// Integer literal exponents lower to OP_IPOW (left-to-right binary powering), constant
// sub-trees made of + - * / sqrt and integer powers are folded with host IEEE arithmetic.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "dtwa_device.cuh"

namespace dtwa {

struct Program {
    std::vector<VmIns> ins;
    std::vector<double> consts;
    int depth = 0, max_depth = 0;  // entries on the VM operand stack

    int add_const(double v)
    {
        for (size_t i = 0; i < consts.size(); ++i)
            if (std::memcmp(&consts[i], &v, sizeof(double)) == 0) return (int)i;
        consts.push_back(v);
        return (int)consts.size() - 1;
    }
    void emit(uint8_t op, uint8_t a = 0, uint16_t b = 0, int delta = 0)
    {
        VmIns i;
        i.op = op;
        i.a = a;
        i.b = b;
        ins.push_back(i);
        depth += delta;
        if (depth > max_depth) max_depth = depth;
    }
};

class ExprCompiler {
   public:
    ExprCompiler(const std::vector<std::string> &vars) : vars_(vars) {}

    // appends code that leaves the value of `text` on the VM stack
    void compile(const std::string &text, Program &prog)
    {
        s_ = text;
        pos_ = 0;
        Node *root = parse_expr();
        skip_ws();
        if (pos_ != s_.size()) fail("unexpected '" + s_.substr(pos_, 8) + "'");
        fold(root);
        gen(root, prog);
        free_node(root);
    }

   private:
    struct Node {
        int kind;  // 0 const, 1 var, 2 unary op, 3 binary op, 4 ipow, 5 pdf
        uint8_t op = 0;
        double val = 0;
        int ivar = 0;
        int n = 0;
        Node *l = nullptr, *r = nullptr;
    };
    const std::vector<std::string> &vars_;
    std::string s_;
    size_t pos_ = 0;

    [[noreturn]] void fail(const std::string &m) const
    {
        throw std::runtime_error("expression '" + s_ + "': " + m);
    }
    static void free_node(Node *n)
    {
        if (!n) return;
        free_node(n->l);
        free_node(n->r);
        delete n;
    }
    void skip_ws()
    {
        while (pos_ < s_.size() && std::isspace((unsigned char)s_[pos_])) ++pos_;
    }
    bool eat(char c)
    {
        skip_ws();
        if (pos_ < s_.size() && s_[pos_] == c) {
            ++pos_;
            return true;
        }
        return false;
    }
    static Node *mk_const(double v)
    {
        Node *n = new Node;
        n->kind = 0;
        n->val = v;
        return n;
    }
    static Node *mk_bin(uint8_t op, Node *l, Node *r)
    {
        Node *n = new Node;
        n->kind = 3;
        n->op = op;
        n->l = l;
        n->r = r;
        return n;
    }
    static Node *mk_un(uint8_t op, Node *l)
    {
        Node *n = new Node;
        n->kind = 2;
        n->op = op;
        n->l = l;
        return n;
    }

    Node *parse_expr()
    {
        Node *l = parse_term();
        for (;;) {
            if (eat('+'))
                l = mk_bin(OP_ADD, l, parse_term());
            else if (eat('-'))
                l = mk_bin(OP_SUB, l, parse_term());
            else
                return l;
        }
    }
    Node *parse_term()
    {
        Node *l = parse_unary();
        for (;;) {
            skip_ws();
            if (pos_ + 1 < s_.size() && s_[pos_] == '*' && s_[pos_ + 1] == '*') return l;  // handled in power
            if (eat('*'))
                l = mk_bin(OP_MUL, l, parse_unary());
            else if (eat('/'))
                l = mk_bin(OP_DIV, l, parse_unary());
            else
                return l;
        }
    }
    Node *parse_unary()
    {
        if (eat('-')) return mk_un(OP_NEG, parse_unary());
        if (eat('+')) return parse_unary();
        return parse_power();
    }
    Node *parse_power()
    {
        Node *base = parse_atom();
        skip_ws();
        bool pw = false;
        if (pos_ < s_.size() && s_[pos_] == '^') {
            ++pos_;
            pw = true;
        } else if (pos_ + 1 < s_.size() && s_[pos_] == '*' && s_[pos_ + 1] == '*') {
            pos_ += 2;
            pw = true;
        }
        if (!pw) return base;
        Node *e = parse_unary();
        // integer literal exponent (possibly negated)?
        bool neg = false;
        Node *lit = e;
        if (lit->kind == 2 && lit->op == OP_NEG) {
            neg = true;
            lit = lit->l;
        }
        if (lit->kind == 0 && lit->val == std::floor(lit->val) && std::fabs(lit->val) <= 1024) {
            Node *n = new Node;
            n->kind = 4;
            n->n = (int)(neg ? -lit->val : lit->val);
            n->l = base;
            free_node(e);
            return n;
        }
        return mk_bin(OP_POW, base, e);
    }
    Node *parse_atom()
    {
        skip_ws();
        if (pos_ >= s_.size()) fail("unexpected end");
        const char c = s_[pos_];
        if (c == '(') {
            ++pos_;
            Node *n = parse_expr();
            if (!eat(')')) fail("missing ')'");
            return n;
        }
        if (std::isdigit((unsigned char)c) || c == '.') {
            const char *b = s_.c_str() + pos_;
            char *end = nullptr;
            const double v = std::strtod(b, &end);
            if (end == b) fail("bad number");
            pos_ += (size_t)(end - b);
            Node *n = mk_const(v);
            // juxtaposition: 2Q, 2(Q+1), 2sqrt(Q)
            if (pos_ < s_.size() && (std::isalpha((unsigned char)s_[pos_]) || s_[pos_] == '(' || (unsigned char)s_[pos_] >= 0x80))
                return mk_bin(OP_MUL, n, parse_power());
            return n;
        }
        // identifier: ASCII letters/digits/underscore, or any non-ASCII UTF-8 bytes (e.g. the Greek pi)
        size_t b = pos_;
        while (pos_ < s_.size() && (std::isalnum((unsigned char)s_[pos_]) || s_[pos_] == '_' || (unsigned char)s_[pos_] >= 0x80)) ++pos_;
        if (b == pos_) fail(std::string("unexpected '") + c + "'");
        const std::string name = s_.substr(b, pos_ - b);
        skip_ws();
        if (pos_ < s_.size() && s_[pos_] == '(') {
            ++pos_;
            std::vector<Node *> args;
            if (!eat(')')) {
                do args.push_back(parse_expr());
                while (eat(','));
                if (!eat(')')) fail("missing ')' after arguments of " + name);
            }
            return mk_call(name, args);
        }
        if (name == "pi" || name == "\xcf\x80") return mk_const(3.14159265358979323846);
        for (size_t i = 0; i < vars_.size(); ++i)
            if (vars_[i] == name) {
                Node *n = new Node;
                n->kind = 1;
                n->ivar = (int)i;
                return n;
            }
        std::string vs;
        for (auto &v : vars_) vs += (vs.empty() ? "" : ",") + v;
        fail("unknown variable '" + name + "'. Write it in terms only of [" + vs + "]");
    }
    Node *mk_call(const std::string &f, std::vector<Node *> &a)
    {
        struct F1 { const char *n; uint8_t op; };
        static const F1 f1[] = {{"sqrt", OP_SQRT}, {"sin", OP_SIN}, {"cos", OP_COS}, {"tan", OP_TAN}, {"exp", OP_EXP},
                                {"log", OP_LOG}, {"abs", OP_ABS}, {"acos", OP_ACOS}, {"asin", OP_ASIN}, {"atan", OP_ATAN},
                                {"\xe2\x88\x9a", OP_SQRT}};
        static const F1 f2[] = {{"atan2", OP_ATAN2}, {"min", OP_MIN}, {"max", OP_MAX}, {"pow", OP_POW}};
        if (f == "__pdf__" && a.empty()) {
            Node *n = new Node;
            n->kind = 5;
            return n;
        }
        if (f == "atan" && a.size() == 2) return mk_bin(OP_ATAN2, a[0], a[1]);  // Julia's two-argument atan
        for (auto &e : f1)
            if (f == e.n) {
                if (a.size() != 1) fail(f + " takes one argument");
                return mk_un(e.op, a[0]);
            }
        for (auto &e : f2)
            if (f == e.n) {
                if (a.size() != 2) fail(f + " takes two arguments");
                return mk_bin(e.op, a[0], a[1]);
            }
        fail("unknown function '" + f + "'");
    }

    static double ipow_host(double x, int n)
    {
        if (n == 0) return 1.0;
        const bool neg = n < 0;
        unsigned m = neg ? (unsigned)(-n) : (unsigned)n;
        int top = 31;
        while (!((m >> top) & 1u)) --top;
        volatile double r = x;
        for (int bit = top - 1; bit >= 0; --bit) {
            r = r * r;
            if ((m >> bit) & 1u) r = r * x;
        }
        return neg ? 1.0 / r : r;
    }
    // fold constant sub-trees whose operations are correctly rounded on both host and device
    static void fold(Node *n)
    {
        if (!n) return;
        fold(n->l);
        fold(n->r);
        auto is_c = [](Node *x) { return x && x->kind == 0; };
        if (n->kind == 3 && is_c(n->l) && is_c(n->r) &&
            (n->op == OP_ADD || n->op == OP_SUB || n->op == OP_MUL || n->op == OP_DIV)) {
            volatile double a = n->l->val, b = n->r->val, v;
            v = n->op == OP_ADD ? a + b : n->op == OP_SUB ? a - b : n->op == OP_MUL ? a * b : a / b;
            set_const(n, v);
        } else if (n->kind == 2 && is_c(n->l) && (n->op == OP_NEG || n->op == OP_SQRT)) {
            set_const(n, n->op == OP_NEG ? -n->l->val : std::sqrt(n->l->val));
        } else if (n->kind == 4 && is_c(n->l)) {
            set_const(n, ipow_host(n->l->val, n->n));
        }
    }
    static void set_const(Node *n, double v)
    {
        free_node(n->l);
        free_node(n->r);
        n->l = n->r = nullptr;
        n->kind = 0;
        n->val = v;
    }
    static bool is_pow2(double c)
    {
        if (!(c > 0) || !std::isfinite(c)) return false;
        int e;
        const double m = std::frexp(c, &e);
        return m == 0.5 && e > -1000 && e < 1000;
    }
    // Emits code that leaves the value of n in the accumulator.  `push`: the accumulator holds a live
    // value that must be saved on the operand stack before it is overwritten.
    static void gen(Node *n, Program &p, bool push = false)
    {
        switch (n->kind) {
        case 0: p.emit(push ? OP_PLDC : OP_LDC, 0, (uint16_t)p.add_const(n->val), push ? +1 : 0); break;
        case 1: p.emit(push ? OP_PLDV : OP_LDV, (uint8_t)n->ivar, 0, push ? +1 : 0); break;
        case 5: p.emit(push ? OP_PLDPDF : OP_LDPDF, 0, 0, push ? +1 : 0); break;
        case 2:
            gen(n->l, p, push);
            p.emit(n->op);
            break;
        case 4:
            gen(n->l, p, push);
            if (n->n == 2)
                p.emit(OP_SQR);
            else if (n->n == 3)
                p.emit(OP_CUBE);
            else
                p.emit(OP_IPOW, 0, (uint16_t)(int16_t)n->n);
            break;
        case 3: {
            Node *L = n->l, *R = n->r;
            const uint8_t op = n->op;
            const bool arith = (op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV);
            if (arith && R->kind == 0) {
                gen(L, p, push);
                if (op == OP_DIV && is_pow2(R->val))   // x / 2^k == x * 2^-k exactly
                    p.emit(OP_MULC, 0, (uint16_t)p.add_const(1.0 / R->val));
                else
                    p.emit(op == OP_ADD ? OP_ADDC : op == OP_SUB ? OP_SUBC : op == OP_MUL ? OP_MULC : OP_DIVC, 0,
                           (uint16_t)p.add_const(R->val));
            } else if (arith && R->kind == 1) {
                gen(L, p, push);
                p.emit(op == OP_ADD ? OP_ADDV : op == OP_SUB ? OP_SUBV : op == OP_MUL ? OP_MULV : OP_DIVV, (uint8_t)R->ivar);
            } else if (arith && L->kind == 0) {   // IEEE + and * commute bit for bit
                gen(R, p, push);
                p.emit(op == OP_ADD ? OP_ADDC : op == OP_SUB ? OP_RSUBC : op == OP_MUL ? OP_MULC : OP_RDIVC, 0,
                       (uint16_t)p.add_const(L->val));
            } else if (arith && L->kind == 1) {
                gen(R, p, push);
                p.emit(op == OP_ADD ? OP_ADDV : op == OP_SUB ? OP_RSUBV : op == OP_MUL ? OP_MULV : OP_RDIVV, (uint8_t)L->ivar);
            } else {
                gen(L, p, push);
                gen(R, p, true);
                p.emit(op, 0, 0, -1);
            }
            break;
        }
        }
    }
};

}  // namespace dtwa
