// LOG() for builds of the host mirror that do not link the reference's src/Decoder.cpp (which defines the
// real one, Decoder.cpp:22).  Same output format: "YYYY-mm-dd HH:MM:SS | message".
#include <cstdarg>
#include <cstdio>
#include <ctime>

void LOG(const char *format, ...)
{
    char log[1024] = {0};
    va_list arg_list;
    va_start(arg_list, format);
    vsnprintf(log, sizeof log, format, arg_list);
    va_end(arg_list);
    time_t cur;
    time(&cur);
    struct tm tmv;
    localtime_r(&cur, &tmv);
    char now[64];
    strftime(now, sizeof now, "%Y-%m-%d %H:%M:%S", &tmv);
    printf("%s | %s\n", now, log);
}
