// parameter_handler.h -- the .prm surface of the reference, without deal.II.
//
// Restates the subset of dealii::ParameterHandler the reference uses
// (declare_parameters cracks.cc:1307-1405, reads at 969-977, 1411-1575,
// 4641-4643): `subsection X ... end`, `set Key = value`, `#` comments anywhere
// on a line (also glued to a value: parameters_miehe_tension_adaptive.prm:17),
// tabs and repeated blanks inside keys, unknown keys / bad values are errors.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cracks {

struct ParameterError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

class ParameterHandler
{
public:
  enum class Pattern { Integer, Double, Bool, Selection, Anything };

  void enter_subsection (const std::string &name);
  void leave_subsection ();
  void declare_entry (const std::string &key, const std::string &default_value, Pattern pattern,
                      const std::string &selection = "", double lower_bound = -1e300);
  void parse_input (const std::string &filename);
  void parse_input_from_string (const std::string &text, const std::string &origin = "<string>");

  std::string get (const std::string &key) const;
  long get_integer (const std::string &key) const;
  double get_double (const std::string &key) const;
  bool get_bool (const std::string &key) const;
  void set (const std::string &key, const std::string &value);
  // ParameterHandler::print_parameters(out, Text)
  std::string print_parameters () const;

private:
  struct Entry
  {
    std::string value, default_value, selection;
    Pattern pattern;
    double lower_bound;
  };
  std::string path () const;
  static std::string collapse (const std::string &s);
  void check (const std::string &key, const Entry &e, const std::string &value, const std::string &where) const;
  std::vector<std::string> current_;
  std::map<std::string, std::map<std::string, Entry>> sections_; // section path -> key -> entry
  std::vector<std::string> section_order_;
};

} // namespace cracks
