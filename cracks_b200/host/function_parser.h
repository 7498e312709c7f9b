// function_parser.h -- one-variable expression evaluator standing in for
// dealii::FunctionParser<1> (muparser) as the reference uses it: `Pressure`
// is a function of `time` (cracks.cc:1490-1491), `K reg` / `Eps reg` are
// functions of `h` (cracks.cc:3876-3882).  Grammar: + - * / ^, unary minus,
// parentheses, numbers in C notation, the variable, pi, and the functions
// pow sqrt sin cos tan exp log abs min max.
#pragma once
#include <stdexcept>
#include <string>

namespace cracks {

class FunctionParser
{
public:
  void initialize (const std::string &variable, const std::string &expression);
  double value (double x) const;

private:
  std::string var_, expr_;
};

} // namespace cracks
