#include "function_parser.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace cracks {

namespace {

struct Parser
{
  const std::string &s;
  const std::string &var;
  double x;
  size_t pos = 0;

  [[noreturn]] void fail (const std::string &why) const
  {
    throw std::invalid_argument ("FunctionParser: " + why + " in <" + s + "> at position " + std::to_string (pos));
  }
  void skip ()
  {
    while (pos < s.size () && std::isspace ((unsigned char) s[pos]))
      ++pos;
  }
  bool eat (char c)
  {
    skip ();
    if (pos < s.size () && s[pos] == c)
      {
        ++pos;
        return true;
      }
    return false;
  }
  double expr ()
  {
    double v = term ();
    for (;;)
      {
        if (eat ('+'))
          v += term ();
        else if (eat ('-'))
          v -= term ();
        else
          return v;
      }
  }
  double term ()
  {
    double v = unary ();
    for (;;)
      {
        if (eat ('*'))
          v *= unary ();
        else if (eat ('/'))
          v /= unary ();
        else
          return v;
      }
  }
  double unary ()
  {
    if (eat ('-'))
      return -unary ();
    if (eat ('+'))
      return unary ();
    return power ();
  }
  double power ()
  {
    const double base = primary ();
    if (eat ('^'))
      return std::pow (base, unary ()); // right associative
    return base;
  }
  double primary ()
  {
    skip ();
    if (pos >= s.size ())
      fail ("unexpected end");
    if (eat ('('))
      {
        const double v = expr ();
        if (!eat (')'))
          fail ("missing ')'");
        return v;
      }
    const char c = s[pos];
    if (std::isdigit ((unsigned char) c) || c == '.')
      {
        char *end = nullptr;
        const double v = std::strtod (s.c_str () + pos, &end);
        if (end == s.c_str () + pos)
          fail ("bad number");
        pos = (size_t) (end - s.c_str ());
        return v;
      }
    if (std::isalpha ((unsigned char) c) || c == '_')
      {
        size_t e = pos;
        while (e < s.size () && (std::isalnum ((unsigned char) s[e]) || s[e] == '_'))
          ++e;
        const std::string name = s.substr (pos, e - pos);
        pos = e;
        if (name == var)
          return x;
        if (name == "pi" || name == "Pi" || name == "PI")
          return 3.14159265358979323846;
        std::vector<double> args;
        if (!eat ('('))
          fail ("unknown identifier <" + name + ">");
        if (!eat (')'))
          {
            do
              args.push_back (expr ());
            while (eat (','));
            if (!eat (')'))
              fail ("missing ')' after arguments of " + name);
          }
        auto need = [&](size_t n) {
          if (args.size () != n)
            fail (name + " expects " + std::to_string (n) + " argument(s)");
        };
        if (name == "pow") { need (2); return std::pow (args[0], args[1]); }
        if (name == "min") { need (2); return std::fmin (args[0], args[1]); }
        if (name == "max") { need (2); return std::fmax (args[0], args[1]); }
        need (1);
        if (name == "sqrt") return std::sqrt (args[0]);
        if (name == "sin") return std::sin (args[0]);
        if (name == "cos") return std::cos (args[0]);
        if (name == "tan") return std::tan (args[0]);
        if (name == "exp") return std::exp (args[0]);
        if (name == "log") return std::log (args[0]);
        if (name == "abs") return std::fabs (args[0]);
        fail ("unknown function <" + name + ">");
      }
    fail (std::string ("unexpected character '") + c + "'");
  }
};

} // namespace

void
FunctionParser::initialize (const std::string &variable, const std::string &expression)
{
  var_ = variable;
  expr_ = expression;
  value (1.0); // validate now, like FunctionParser::initialize does
}

double
FunctionParser::value (double x) const
{
  Parser p{expr_, var_, x};
  const double v = p.expr ();
  p.skip ();
  if (p.pos != expr_.size ())
    p.fail ("trailing characters");
  return v;
}

} // namespace cracks
