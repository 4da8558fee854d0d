/* JNI entry of SparkBWA's bridge: Java_com_github_sparkbwa_BwaJni_bwa_1jni
 * (declared in src/main/native/com_github_sparkbwa_BwaJni.h:15-16, implemented by
 * src/main/native/bwa_jni.c:29-158, bound by java/com/github/sparkbwa/BwaJni.java:59 with
 * descriptor (I[Ljava/lang/String;[I)I).
 *
 * Behaviour kept: argv comes from the Java String[]; "-f <file>" names the SAM output and is
 * removed from the argument list for `mem` (bwa_jni.c:74-111); the return value is the aligner's
 * return code.  Behaviour changed on purpose: the SAM text is written straight to
 * the -f path instead of freopen()-ing the process-wide stdout, the UTF strings are released, and
 * nothing exits the JVM.
 *
 * No jni.h exists in the build image, so the few JNI types and the four function-table slots used
 * here are declared with their ABI-stable indices (JNI spec, "Interface Function Table":
 * GetStringUTFChars = 169, ReleaseStringUTFChars = 170, GetArrayLength = 171,
 * GetObjectArrayElement = 173).  With a real jni.h present, compile with -DB2_HAVE_JNI_H.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/b200bwa.h"

#ifdef B2_HAVE_JNI_H
#include <jni.h>
#else
typedef int jint;
typedef int jsize;
typedef unsigned char jboolean;
typedef void* jobject;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jintArray;
typedef const void* const* JNIEnv; /* JNIEnv = pointer to the function table */
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef const char* (*b2_GetStringUTFChars_t)(JNIEnv*, jstring, jboolean*);
typedef void (*b2_ReleaseStringUTFChars_t)(JNIEnv*, jstring, const char*);
typedef jsize (*b2_GetArrayLength_t)(JNIEnv*, jarray);
typedef jobject (*b2_GetObjectArrayElement_t)(JNIEnv*, jobjectArray, jsize);
#endif

JNIEXPORT jint JNICALL Java_com_github_sparkbwa_BwaJni_bwa_1jni(JNIEnv* env, jobject thisObj, jint argc,
                                                                 jobjectArray stringArray, jintArray lenStrings) {
  (void)thisObj; (void)lenStrings; (void)argc;
#ifdef B2_HAVE_JNI_H
  jsize n = (*env)->GetArrayLength(env, stringArray);
#else
  const void* const* tab = *env;
  b2_GetStringUTFChars_t GetStringUTFChars = (b2_GetStringUTFChars_t)tab[169];
  b2_ReleaseStringUTFChars_t ReleaseStringUTFChars = (b2_ReleaseStringUTFChars_t)tab[170];
  b2_GetArrayLength_t GetArrayLength = (b2_GetArrayLength_t)tab[171];
  b2_GetObjectArrayElement_t GetObjectArrayElement = (b2_GetObjectArrayElement_t)tab[173];
  jsize n = GetArrayLength(env, stringArray);
#endif
  if (n <= 0) return 1;
  jstring* js = (jstring*)calloc((size_t)n, sizeof(jstring));
  const char** utf = (const char**)calloc((size_t)n, sizeof(char*));
  char** argv = (char**)calloc((size_t)n + 1, sizeof(char*));
  const char* out = 0;
  int i, na = 0, ret = 1, bad = 0;
  if (!js || !utf || !argv) { free(js); free(utf); free(argv); return 1; }
  for (i = 0; i < n && !bad; ++i) {
    /* a null String element, or a NULL from GetStringUTFChars (out of memory, exception pending): give up with
     * rc 1 after releasing what was acquired -- nothing may crash the executor JVM */
#ifdef B2_HAVE_JNI_H
    js[i] = (jstring)(*env)->GetObjectArrayElement(env, stringArray, i);
    utf[i] = js[i] ? (*env)->GetStringUTFChars(env, js[i], 0) : 0;
#else
    js[i] = (jstring)GetObjectArrayElement(env, stringArray, i);
    utf[i] = js[i] ? GetStringUTFChars(env, js[i], 0) : 0;
#endif
    if (!utf[i]) bad = 1;
  }
  if (bad) {
    fprintf(stderr, "[%s] null or unreadable argument string; nothing run\n", __func__);
    goto release;
  }
  for (i = 0; i < n; ++i) {
    if (strcmp(utf[i], "-f") == 0 && i < n - 1) { out = utf[++i]; continue; } /* output file, not an aligner option */
    argv[na++] = (char*)utf[i];
  }
  fprintf(stderr, "[%s] %d arguments, output '%s'\n", __func__, na, out ? out : "(stdout)");
  ret = b200bwa_main_to(na, argv, out);
  fprintf(stderr, "[%s] Return code from BWA %d.\n", __func__, ret);
release:
  for (i = 0; i < n; ++i) {
    if (!utf[i]) continue;
#ifdef B2_HAVE_JNI_H
    (*env)->ReleaseStringUTFChars(env, js[i], utf[i]);
#else
    ReleaseStringUTFChars(env, js[i], utf[i]);
#endif
  }
  free(js); free(utf); free(argv);
  return ret;
}
