#include "parameter_handler.h"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace cracks {

std::string
ParameterHandler::collapse (const std::string &s)
{
  // trim and replace every run of blanks/tabs by one blank
  std::string out;
  bool in_space = true;
  for (char c : s)
    {
      const bool sp = c == ' ' || c == '\t' || c == '\r' || c == '\n';
      if (sp)
        {
          if (!in_space)
            out.push_back (' ');
          in_space = true;
        }
      else
        {
          out.push_back (c);
          in_space = false;
        }
    }
  while (!out.empty () && out.back () == ' ')
    out.pop_back ();
  return out;
}

std::string
ParameterHandler::path () const
{
  std::string p;
  for (const auto &s : current_)
    p += "/" + s;
  return p;
}

void
ParameterHandler::enter_subsection (const std::string &name)
{
  current_.push_back (collapse (name));
  const std::string p = path ();
  if (!sections_.count (p))
    {
      sections_[p];
      section_order_.push_back (p);
    }
}

void
ParameterHandler::leave_subsection ()
{
  if (current_.empty ())
    throw ParameterError ("leave_subsection without enter_subsection");
  current_.pop_back ();
}

void
ParameterHandler::declare_entry (const std::string &key, const std::string &default_value, Pattern pattern,
                                 const std::string &selection, double lower_bound)
{
  const std::string p = path ();
  if (!sections_.count (p))
    {
      sections_[p];
      section_order_.push_back (p);
    }
  Entry e{default_value, default_value, selection, pattern, lower_bound};
  check (key, e, default_value, "default");
  sections_[p][collapse (key)] = e;
}

void
ParameterHandler::check (const std::string &key, const Entry &e, const std::string &value,
                         const std::string &where) const
{
  auto bad = [&](const std::string &why) {
    throw ParameterError (where + ": value <" + value + "> of entry <" + key + "> " + why);
  };
  char *end = nullptr;
  switch (e.pattern)
    {
    case Pattern::Integer:
      {
        const long v = std::strtol (value.c_str (), &end, 10);
        if (value.empty () || *end != '\0')
          bad ("is not an integer");
        if ((double) v < e.lower_bound)
          bad ("is below the lower bound");
        break;
      }
    case Pattern::Double:
      {
        const double v = std::strtod (value.c_str (), &end);
        if (value.empty () || *end != '\0')
          bad ("is not a floating point number");
        if (v < e.lower_bound)
          bad ("is below the lower bound");
        break;
      }
    case Pattern::Bool:
      if (value != "true" && value != "false")
        bad ("is not 'true' or 'false'");
      break;
    case Pattern::Selection:
      {
        std::stringstream ss (e.selection);
        std::string item;
        bool ok = false;
        while (std::getline (ss, item, '|'))
          ok = ok || item == value;
        if (!ok)
          bad ("is not one of <" + e.selection + ">");
        break;
      }
    case Pattern::Anything:
      break;
    }
}

void
ParameterHandler::parse_input (const std::string &filename)
{
  std::ifstream in (filename.c_str ());
  if (!in)
    throw ParameterError ("cannot open parameter file <" + filename + ">");
  std::stringstream ss;
  ss << in.rdbuf ();
  parse_input_from_string (ss.str (), filename);
}

void
ParameterHandler::parse_input_from_string (const std::string &text, const std::string &origin)
{
  const std::vector<std::string> saved = current_;
  current_.clear ();
  std::stringstream ss (text);
  std::string line;
  int lineno = 0;
  while (std::getline (ss, line))
    {
      ++lineno;
      const std::string where = origin + ":" + std::to_string (lineno);
      const size_t hash = line.find ('#');
      if (hash != std::string::npos)
        line.erase (hash);
      line = collapse (line);
      if (line.empty ())
        continue;
      if (line.compare (0, 11, "subsection ") == 0)
        {
          const std::string name = collapse (line.substr (11));
          current_.push_back (name);
          if (!sections_.count (path ()))
            throw ParameterError (where + ": no such subsection <" + name + ">");
        }
      else if (line == "end")
        {
          if (current_.empty ())
            throw ParameterError (where + ": 'end' without subsection");
          current_.pop_back ();
        }
      else if (line.compare (0, 4, "set ") == 0)
        {
          const size_t eq = line.find ('=');
          if (eq == std::string::npos)
            throw ParameterError (where + ": missing '=' in <" + line + ">");
          const std::string key = collapse (line.substr (4, eq - 4));
          const std::string value = collapse (line.substr (eq + 1));
          auto &sec = sections_[path ()];
          auto it = sec.find (key);
          if (it == sec.end ())
            throw ParameterError (where + ": no entry with name <" + key + "> was declared in the current subsection");
          check (key, it->second, value, where);
          it->second.value = value;
        }
      else
        throw ParameterError (where + ": cannot parse <" + line + ">");
    }
  if (!current_.empty ())
    throw ParameterError (origin + ": unbalanced 'subsection'/'end'");
  current_ = saved;
}

std::string
ParameterHandler::get (const std::string &key) const
{
  auto s = sections_.find (path ());
  if (s == sections_.end ())
    throw ParameterError ("no such subsection " + path ());
  auto it = s->second.find (collapse (key));
  if (it == s->second.end ())
    throw ParameterError ("no entry <" + key + "> in " + path ());
  return it->second.value;
}

long
ParameterHandler::get_integer (const std::string &key) const
{
  return std::strtol (get (key).c_str (), nullptr, 10);
}

double
ParameterHandler::get_double (const std::string &key) const
{
  return std::strtod (get (key).c_str (), nullptr);
}

bool
ParameterHandler::get_bool (const std::string &key) const
{
  return get (key) == "true";
}

void
ParameterHandler::set (const std::string &key, const std::string &value)
{
  auto &sec = sections_[path ()];
  auto it = sec.find (collapse (key));
  if (it == sec.end ())
    throw ParameterError ("no entry <" + key + "> in " + path ());
  check (key, it->second, value, "set");
  it->second.value = value;
}

std::string
ParameterHandler::print_parameters () const
{
  std::stringstream out;
  out << "# Listing of Parameters\n# ---------------------\n";
  for (const auto &p : section_order_)
    {
      const auto &sec = sections_.at (p);
      const bool top = p.empty ();
      if (!top)
        out << "subsection " << p.substr (1) << "\n";
      for (const auto &kv : sec)
        out << (top ? "" : "  ") << "set " << kv.first << " = " << kv.second.value << "\n";
      if (!top)
        out << "end\n\n";
    }
  return out.str ();
}

} // namespace cracks
