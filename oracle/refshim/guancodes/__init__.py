"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored PyPI package ``guancodes`` (pinned
``~=0.0.3`` in the reference's requirements.txt:4), so that the *reference* imports in this
container.  Never imported by the product package ``theboss_b200``."""
